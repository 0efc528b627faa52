"""Diagnostic: C4-shaped solve (n^2 variable-coefficient operator) with the solver's verbose trace,
scaled vs classic Jacobi-PCG.   python profiles/diag_c4.py [n=2048]"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch  # noqa: E402

import bench_configs as bc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
dev = torch.device("cuda")
h_ = 1.0 / (n + 1)
xs = torch.arange(0, n + 2, device=dev, dtype=torch.float64) * h_
X, Y = torch.meshgrid(xs, xs, indexing="ij")
kext = (1 + X ** 2 + Y ** 2).contiguous()
Xi, Yi = X[1:-1, 1:-1], Y[1:-1, 1:-1]
rhs = (-2 * Xi * (1 - Xi) * (Xi ** 2 + Yi ** 2 + 1) + 2 * Xi * (-Xi * Yi * (1 - Yi) + Yi * (1 - Xi) * (1 - Yi))
       - 2 * Yi * (1 - Yi) * (Xi ** 2 + Yi ** 2 + 1) + 2 * Yi * (-Xi * Yi * (1 - Xi) + Xi * (1 - Xi) * (1 - Yi)))
f = (-(2 * h_ * h_) * rhs).reshape(-1).contiguous()
h = C.c_int64(0)
bc.pkg.check(bc.lib.ab_poisson2d_csr(n, kext.data_ptr(), -1, C.byref(h)))
u = torch.empty_like(f)
bc.lib.ab_set_option(b"verbose", 1.0)
for scaled in (1.0, 0.0):
    bc.lib.ab_set_option(b"cg_scaled", scaled)
    it, rr = C.c_int(0), C.c_double(0)
    t0 = time.perf_counter()
    bc.pkg.check(bc.lib.ab_solve(h.value, b"CG", f.data_ptr(), u.data_ptr(), 1e-10, 200000, C.byref(it), C.byref(rr)))
    torch.cuda.synchronize()
    print(f"cg_scaled={scaled}: iters {it.value} relres {rr.value:.3e} seconds {time.perf_counter() - t0:.3f}", flush=True)
