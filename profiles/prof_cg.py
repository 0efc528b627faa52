"""Short CG run for ncu (launch list + one full capture of the SpMV kernel).

    python profiles/prof_cg.py [grid=512] [iters=8]

Builds the n^3 7-point Poisson CSR on the device and runs `iters` Jacobi-PCG iterations through
the C-ABI (ab_cg_fixed) — the same launches bench.py times, minus the Python-side timing."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 8
pkg = ge.load_package()
for kv in filter(None, os.environ.get("AB_OPTS", "").split(",")):      # A/B switches, e.g. AB_OPTS=spmv_zmarch_g=2
    k_, v_ = kv.split("=")
    pkg.check(pkg.lib.ab_set_option(k_.strip().encode(), float(v_)))
h = C.c_int64(0)
pkg.check(pkg.lib.ab_poisson3d_csr(n, n, n, 0, n ** 3, C.byref(h)))
A = pkg.SparseTensor.from_handle(h.value)
b = A @ torch.ones(n ** 3, dtype=torch.float64, device="cuda")
u = torch.empty_like(b)
rel = C.c_double(0)
pkg.check(pkg.lib.ab_cg_fixed(h.value, b.data_ptr(), u.data_ptr(), iters, C.byref(rel)))
torch.cuda.synchronize()
print(f"grid {n}^3, {iters} CG iterations, relres {rel.value:.6e}, launches {pkg.lib.ab_launch_count()}")
