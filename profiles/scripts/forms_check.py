"""Bit-compare every solver SpMV form against the plain int32 kernel on n^3 Poisson grids given on the command line."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import __graft_entry__ as ge
pkg = ge.load_package(); lib = pkg.lib
keys = (b"spmv_rowpat", b"spmv_code", b"spmv_off8", b"spmv_rowwin", b"spmv_rowmarch", b"spmv_zmarch")
forms = (("zmarch", 1, 1, 1, 1, 1, 1), ("rowmarch", 1, 1, 1, 1, 1, 0), ("rowwin", 1, 1, 1, 1, 0, 0), ("rowpat", 1, 1, 1, 0, 0, 0),
         ("code", 0, 1, 1, 0, 0, 0), ("off8", 0, 0, 1, 0, 0, 0), ("plain", 0, 0, 0, 0, 0, 0))
def info(h, k):
    v = C.c_double(0); lib.ab_csr_get_info(h, k, C.byref(v)); return v.value
for n in (int(a) for a in sys.argv[1:]):
    N = n ** 3
    h = C.c_int64(0)
    pkg.check(lib.ab_poisson3d_csr(n, n, n, 0, N, C.byref(h)))
    A = pkg.SparseTensor.from_handle(h.value)
    g = torch.Generator(device="cuda").manual_seed(233)
    x = torch.randn(N, dtype=torch.float64, device="cuda", generator=g)
    ys = {}
    for name, *vals in forms:
        for k, v in zip(keys, vals): pkg.check(lib.ab_set_option(k, float(v)))
        y = torch.empty_like(x)
        pkg.check(lib.ab_spmv_solver_form(A.handle(), x.data_ptr(), y.data_ptr()))
        torch.cuda.synchronize()
        ys[name] = y
    for k in keys: pkg.check(lib.ab_set_option(k, 1.0))
    res = {name: (bool(torch.equal(y, ys["plain"])), float((y - ys["plain"]).abs().max()), bool(torch.isfinite(y).all())) for name, y in ys.items()}
    print(n, {k: info(h.value, k) for k in (b"rowwin", b"rowmarch", b"zmarch", b"zmarch_g")}, res, flush=True)
    A.close()
