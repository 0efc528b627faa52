import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package(); lib = pkg.lib
rng = np.random.default_rng(233)
n = 64
f = rng.uniform(size=n * n)
def info(h, key):
    v = C.c_double(0); lib.ab_csr_get_info(h, key, C.byref(v)); return v.value
for opts in ({}, {"spmv_rowwin": 0.0}, {"spmv_rowpat": 0.0}):
    for k, v in opts.items(): pkg.check(lib.ab_set_option(k.encode(), v))
    h2 = C.c_int64(0)
    k1, k3 = np.ones((n + 2, n + 2)), 3.0 * np.ones((n + 2, n + 2))
    pkg.check(lib.ab_poisson2d_csr(n, k1.ctypes.data, -1, C.byref(h2)))
    B = pkg.SparseTensor.from_handle(h2.value)
    S1 = B.to_scipy()
    u1 = pkg.backslash(B, f, "CG")
    r1 = np.linalg.norm(S1 @ u1 - f) / np.linalg.norm(f)
    i1 = [info(h2.value, k) for k in (b"npat", b"ncode", b"rowwin", b"rowmarch", b"scaled")]
    pkg.check(lib.ab_poisson2d_update(h2.value, n, k3.ctypes.data, -1))
    S3 = B.to_scipy()
    u3 = pkg.backslash(B, f, "CG")
    r3 = np.linalg.norm(S3 @ u3 - f) / np.linalg.norm(f)
    i3 = [info(h2.value, k) for k in (b"npat", b"ncode", b"rowwin", b"rowmarch", b"scaled")]
    u3b = pkg.backslash(B, f, "CG")
    r3b = np.linalg.norm(S3 @ u3b - f) / np.linalg.norm(f)
    print(opts, "res1 %.2e res3 %.2e res3b %.2e" % (r1, r3, r3b), "diff %.2e" % (np.linalg.norm(3 * u3 - u1) / np.linalg.norm(u1)), i1, i3,
          "S3/S1 diag", S3.diagonal()[:2], S1.diagonal()[:2], flush=True)
    for k in opts: pkg.check(lib.ab_set_option(k.encode(), 1.0))
    B.close()
