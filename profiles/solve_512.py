"""Time-to-solution of the BASELINE system: Jacobi-PCG on the 3-D 7-point Poisson n^3 system,
b = A*1, x0 = 0, solved to a relative residual of 1e-10 through ab_solve (device-resident vectors).
    python profiles/solve_512.py gpurun_out/solve_512.json [n=512]"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/solve_512.json"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 512
N = n ** 3
pkg = ge.load_package()
lib = pkg.lib
h = C.c_int64(0)
pkg.check(lib.ab_poisson3d_csr(n, n, n, 0, N, C.byref(h)))
A = pkg.SparseTensor.from_handle(h.value)
ones = torch.ones(N, dtype=torch.float64, device="cuda")
b = A @ ones
u = torch.empty_like(b)
it, rr = C.c_int(0), C.c_double(0)
out = {"system": f"3-D 7-point Poisson {n}^3, b = A*1, x0 = 0, tol 1e-10", "rows": N}
for rep in range(2):                       # the first solve also builds the dictionary forms
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pkg.check(lib.ab_solve(h.value, b"CG", b.data_ptr(), u.data_ptr(), 1e-10, 200000, C.byref(it), C.byref(rr)))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    true_rel = float(((A @ u) - b).norm() / b.norm())
    out["first_solve" if rep == 0 else "solve"] = {"seconds": dt, "iterations": it.value, "iters_per_s": it.value / dt,
                                                  "reported_relres": rr.value, "true_relres": true_rel,
                                                  "max_err_vs_ones": float((u - 1).abs().max())}
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out), flush=True)
