"""SpMV kernel variants on the n^3 7-point Poisson matrix (CUDA-event timing through ab_bench_spmv).

    python profiles/bench_spmv_variants.py [grid=512]
"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402

import bench_configs as bc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = n ** 3
h = C.c_int64(0)
bc.pkg.check(bc.lib.ab_poisson3d_csr(n, n, n, 0, N, C.byref(h)))
nnz = 7 * N - 6 * n * n
alg = 12.0 * nnz + 4.0 * (N + 1) + 16.0 * N
x = torch.rand(N, dtype=torch.float64, device="cuda") * 2 - 1
y = torch.empty_like(x)
ref = None
cases = [("int32 colind", dict(spmv_off8=0)),
         ("off8 dictionary", dict(spmv_off8=1)),
         ("off8 dictionary, U=4", dict(spmv_off8=1, spmv_variant=2)),
         ("off8 dictionary, 6 CTAs/SM grid", dict(spmv_off8=1, spmv_ctas_per_sm=6))]
out = {}
for name, opts in cases:
    bc.lib.ab_set_option(b"spmv_variant", -1.0)
    for k, v in opts.items():
        bc.lib.ab_set_option(k.encode(), float(v))
    t = C.c_double(0)
    for with_dot in (0, 1):
        bc.pkg.check(bc.lib.ab_bench_spmv(h.value, x.data_ptr(), y.data_ptr(), 5, with_dot, C.byref(t)))
        bc.pkg.check(bc.lib.ab_bench_spmv(h.value, x.data_ptr(), y.data_ptr(), 30, with_dot, C.byref(t)))
        out[f"{name} dot={with_dot}"] = {"ms": t.value, "alg_GBps": alg / t.value / 1e6}
    if ref is None:
        ref = y.clone()
    out[name + " max|diff|"] = float((y - ref).abs().max())
print(json.dumps(out, indent=1))
