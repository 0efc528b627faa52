"""A few launches of the offset-dictionary SpMV kernels (plain and pipelined) for ncu.
    python profiles/prof_spmv8.py [grid=512]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402

import bench_configs as bc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = n ** 3
h = C.c_int64(0)
bc.pkg.check(bc.lib.ab_poisson3d_csr(n, n, n, 0, N, C.byref(h)))
x = torch.rand(N, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
t = C.c_double(0)
for pipe in (0.0, 1.0):
    bc.lib.ab_set_option(b"spmv_pipe", pipe)
    bc.pkg.check(bc.lib.ab_bench_spmv(h.value, x.data_ptr(), y.data_ptr(), 3, 1, C.byref(t)))
    print("pipe", pipe, "ms", t.value)
