"""Measurements for the widened rows: structural ops, SpGEMM, A+B, Newton step on a
2-D 5-point FEM-style problem, CUDA events on the library's stream, with scipy on one host core beside
them (the reference ops are single-threaded Eigen/STL loops).  Writes one JSON object.

    python profiles/bench_structure.py gpurun_out/structure.json [n=2048]
"""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import scipy.sparse as sp  # noqa: E402
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_package()
lib = pkg.lib
dev = torch.device("cuda")
PEAK = 6544.3
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def ev(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def host(fn, reps=2):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts)


def facewise_2d_device(n, seed=233):
    idx = torch.arange(n * n, device=dev, dtype=torch.int64).view(n, n)
    I, J, V = [], [], []

    def faces(a, b):
        a, b = a.reshape(-1), b.reshape(-1)
        one = torch.ones(a.numel(), dtype=torch.float64, device=dev)
        I.extend([a, b, a, b]); J.extend([a, b, b, a]); V.extend([one, one, -one, -one])

    faces(idx[:, :-1], idx[:, 1:]); faces(idx[:-1, :], idx[1:, :])
    for edge in (idx[0], idx[-1], idx[:, 0], idx[:, -1]):
        e = edge.reshape(-1)
        I.append(e); J.append(e); V.append(torch.ones(e.numel(), dtype=torch.float64, device=dev))
    ii, jj, vv = torch.cat(I) + 1, torch.cat(J) + 1, torch.cat(V)
    p = torch.randperm(ii.numel(), device=dev, generator=torch.Generator(device=dev).manual_seed(seed))
    return ii[p].contiguous(), jj[p].contiguous(), vv[p].contiguous()


def trip_fetch(th):
    k = lib.ab_trip_size(th)
    oi = torch.empty(k, dtype=torch.int64, device=dev); oj = torch.empty_like(oi)
    ov = torch.empty(k, dtype=torch.float64, device=dev)
    pkg.check(lib.ab_trip_fetch(th, oi.data_ptr(), oj.data_ptr(), ov.data_ptr()))
    return oi, oj, ov


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/structure.json"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    N = n * n
    out = {"peak_GBps": PEAK, "problem": f"2-D 5-point face-wise stream on {n}x{n} (N = {N}), seed-233 shuffle"}
    ii, jj, vv = facewise_2d_device(n)
    T = ii.numel()
    out["T"] = T
    grid = torch.arange(N, device=dev, dtype=torch.int64).view(n, n)
    bd = torch.unique(torch.cat([grid[0], grid[-1], grid[:, 0], grid[:, -1]])) + 1           # boundary DoFs, 1-based
    out["nbd"] = int(bd.numel())
    # host copies for the scipy leg
    hi, hj, hv = (ii - 1).cpu().numpy(), (jj - 1).cpu().numpy(), vv.cpu().numpy()
    A0 = sp.coo_matrix((hv, (hi, hj)), shape=(N, N)).tocsr()
    hbd = (bd - 1).cpu().numpy()

    # ---- zero_out_row (ZeroOutRow.h:28-57): indices are the 0-based [T,2] matrix
    ind = torch.stack([ii - 1, jj - 1], dim=1).contiguous()
    res = {}

    def zor():
        th = C.c_int64(0)
        pkg.check(lib.ab_zero_out_row(T, ind.data_ptr(), vv.data_ptr(), bd.data_ptr(), bd.numel(), C.byref(th)))
        res["zor"] = trip_fetch(th.value)

    ms = ev(zor)
    Tout = res["zor"][2].numel()
    byt = 24.0 * T + 24.0 * Tout
    keep = ~np.isin(hi, hbd)

    def zor_cpu():
        m_ = ~np.isin(hi, hbd)
        return hi[m_], hj[m_], hv[m_]

    out["zero_out_row"] = {"ms": ms, "T_in": T, "T_out": Tout, "algorithmic_GB": byt / 1e9, "GBps": byt / ms / 1e6,
                           "frac_of_peak": byt / ms / 1e6 / PEAK, "check_count": int(keep.sum()) == Tout,
                           "cpu_numpy_ms": host(zor_cpu)}
    del ind

    # ---- scatter_update (SparseScatterUpdate.h:20-60): replace the block of the first 1000 DoFs
    nb = 1000
    sel = torch.arange(1, nb + 1, device=dev, dtype=torch.int64)
    Bm = sp.random(nb, nb, 0.01, random_state=7, format="coo")
    ui = torch.tensor(Bm.row + 1, device=dev, dtype=torch.int64); uj = torch.tensor(Bm.col + 1, device=dev, dtype=torch.int64)
    uv = torch.tensor(Bm.data, device=dev)

    def ssu():
        th = C.c_int64(0)
        pkg.check(lib.ab_scatter_update(T, ii.data_ptr(), jj.data_ptr(), vv.data_ptr(), ui.numel(), ui.data_ptr(), uj.data_ptr(),
                                        uv.data_ptr(), N, N, sel.data_ptr(), nb, sel.data_ptr(), nb, C.byref(th)))
        res["ssu"] = trip_fetch(th.value)

    ms = ev(ssu)
    Tout = res["ssu"][2].numel()
    byt = 24.0 * T + 24.0 * Tout
    out["scatter_update"] = {"ms": ms, "T_out": Tout, "algorithmic_GB": byt / 1e9, "GBps": byt / ms / 1e6,
                             "frac_of_peak": byt / ms / 1e6 / PEAK}

    # ---- assembly + getindex A[interior, interior] (SparseIndexing.h:23-83) on the assembled handle
    h = C.c_int64(0)
    ms_asm = ev(lambda: (lib.ab_csr_destroy(h.value) if h.value else 0,
                         pkg.check(lib.ab_csr_from_coo(T, ii.data_ptr(), jj.data_ptr(), vv.data_ptr(), N, N, 1, C.byref(h)))), reps=2, warm=1)
    A = pkg.SparseTensor.from_handle(h.value)
    nnz = A.query()[2]
    out["assembly"] = {"ms": ms_asm, "nnz": nnz, "Mtriplets_per_s": T / ms_asm / 1e3}
    mask = torch.ones(N, dtype=torch.bool, device=dev); mask[bd - 1] = False
    interior = (torch.nonzero(mask).reshape(-1) + 1).contiguous()
    ni = interior.numel()

    def sidx():
        th = C.c_int64(0)
        pkg.check(lib.ab_sparse_indexing(h.value, interior.data_ptr(), ni, interior.data_ptr(), ni, C.byref(th)))
        res["sidx"] = trip_fetch(th.value)

    ms = ev(sidx)
    nout = res["sidx"][2].numel()
    hint = (interior - 1).cpu().numpy()
    S0 = None

    def sidx_cpu():
        nonlocal S0
        S0 = A0[hint][:, hint]

    cpu_ms = host(sidx_cpu, reps=1)
    out["getindex"] = {"ms": ms, "ni": ni, "nnz_out": nout, "check_count": int(S0.nnz) == nout, "Mentries_per_s": nnz / ms / 1e3,
                       "cpu_scipy_ms": cpu_ms}

    # ---- A + B and SpGEMM A*A on device handles
    hc = C.c_int64(0)

    def axpby():
        if hc.value:
            lib.ab_csr_destroy(hc.value)
        pkg.check(lib.ab_csr_axpby(h.value, 1.0, h.value, 2.0, C.byref(hc)))

    ms = ev(axpby)
    out["add"] = {"ms": ms, "Mentries_per_s": 2 * nnz / ms / 1e3, "cpu_scipy_ms": host(lambda: A0 + 2.0 * A0)}
    lib.ab_csr_destroy(hc.value); hc = C.c_int64(0)

    def spgemm():
        if hc.value:
            lib.ab_csr_destroy(hc.value)
        pkg.check(lib.ab_spgemm(h.value, h.value, 0, C.byref(hc)))

    ms = ev(spgemm, reps=2)
    Cq = pkg.SparseTensor.from_handle(hc.value).query()
    C0 = None

    def spgemm_cpu():
        nonlocal C0
        C0 = A0 @ A0

    cpu_ms = host(spgemm_cpu, reps=1)
    out["spgemm"] = {"ms": ms, "products": Cq[3], "nnz_C": Cq[2], "Mproducts_per_s": Cq[3] / ms / 1e3,
                     "check_nnz": int(C0.nnz) == Cq[2], "cpu_scipy_ms": cpu_ms}

    # ---- one Newton step: delta = J \\ r, u - delta, ||r||  (J = K + 3 diag(u^2), SPD)
    u = torch.rand(N, dtype=torch.float64, device=dev) + 0.5
    dI = torch.arange(1, N + 1, device=dev, dtype=torch.int64)
    Ji, Jj, Jv = torch.cat([ii, dI]), torch.cat([jj, dI]), torch.cat([vv, 3 * u * u])
    hJ = C.c_int64(0)
    pkg.check(lib.ab_csr_from_coo(Jv.numel(), Ji.data_ptr(), Jj.data_ptr(), Jv.data_ptr(), N, N, 1, C.byref(hJ)))
    r = torch.rand(N, dtype=torch.float64, device=dev)
    unew = torch.empty_like(u); delta = torch.empty_like(u)
    rn, li = C.c_double(0), C.c_int(0)

    def newton():
        pkg.check(lib.ab_newton_step(hJ.value, b"SparseLU", r.data_ptr(), u.data_ptr(), 1.0, unew.data_ptr(), delta.data_ptr(),
                                     C.byref(rn), 1e-10, 0, C.byref(li)))

    ms = ev(newton, reps=2)
    J0 = A0 + sp.diags(3 * u.cpu().numpy() ** 2)
    resid = float(np.linalg.norm(J0 @ delta.cpu().numpy() - r.cpu().numpy()) / np.linalg.norm(r.cpu().numpy()))
    ms_upd = ev(lambda: pkg.check(lib.ab_csr_update_values(hJ.value, Jv.data_ptr())))
    out["newton_step"] = {"ms": ms, "linear_iters": li.value, "true_relres_of_delta": resid, "rnorm": rn.value,
                          "jacobian_value_refresh_ms": ms_upd}
    lib.ab_csr_destroy(hJ.value)
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
