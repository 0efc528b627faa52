"""Secondary measurements on BASELINE.json's other configs (parity-test cases, not bench lines):
C2 assembly + SpMV fwd/grad at 256^3, C4 inverse-conductivity forward + adjoint at 4096^2,
C5 power-law SpMM + value gradient.  Writes one JSON object to the path given as argv[1].

    python profiles/bench_configs.py gpurun_out/configs.json [c2,c4,c5] [scale]

All inputs are generated on the device (torch, seed 233); timings are CUDA events on the library's
stream; every section verifies its result through a size-independent property before timing counts.
"""
import ctypes as C
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

PEAK = 6544.3
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
pkg = ge.load_package()
lib = pkg.lib
dev = torch.device("cuda")


def ev_time(fn, reps=1, warm=0):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def facewise_3d_device(n, seed=233):
    idx = torch.arange(n ** 3, device=dev, dtype=torch.int64).view(n, n, n)
    I, J, V = [], [], []

    def faces(a, b):
        a, b = a.reshape(-1), b.reshape(-1)
        one = torch.ones(a.numel(), dtype=torch.float64, device=dev)
        I.extend([a, b, a, b]); J.extend([a, b, b, a]); V.extend([one, one, -one, -one])

    faces(idx[:, :, :-1], idx[:, :, 1:]); faces(idx[:, :-1, :], idx[:, 1:, :]); faces(idx[:-1, :, :], idx[1:, :, :])
    for face in (idx[0], idx[-1], idx[:, 0], idx[:, -1], idx[:, :, 0], idx[:, :, -1]):
        f = face.reshape(-1)
        I.append(f); J.append(f); V.append(torch.ones(f.numel(), dtype=torch.float64, device=dev))
    ii, jj, vv = torch.cat(I) + 1, torch.cat(J) + 1, torch.cat(V)
    del I, J, V
    g = torch.Generator(device=dev).manual_seed(seed)
    p = torch.randperm(ii.numel(), device=dev, generator=g)
    return ii[p].contiguous(), jj[p].contiguous(), vv[p].contiguous()


def c2(n=256):
    out = {"config": f"C2: 3-D 7-pt Laplacian {n}^3, face-wise triplet stream, seed-233 shuffle"}
    N = n ** 3
    ii, jj, vv = facewise_3d_device(n)
    T = ii.numel()
    out["T"] = T
    h = C.c_int64(0)

    def assemble():
        if h.value:
            pkg.check(lib.ab_csr_destroy(h.value))
        pkg.check(lib.ab_csr_from_coo(T, ii.data_ptr(), jj.data_ptr(), vv.data_ptr(), N, N, 1, C.byref(h)))

    assemble()
    ms = min(ev_time(assemble) for _ in range(3))
    A = pkg.SparseTensor.from_handle(h.value)
    m_, n_, nnz, T_ = A.query()
    out["nnz"] = nnz
    # verify against the direct-to-CSR generator, bit-exact, on the device
    g = C.c_int64(0)
    pkg.check(lib.ab_poisson3d_csr(n, n, n, 0, N, C.byref(g)))
    G = pkg.SparseTensor.from_handle(g.value)

    def export(hh, cnt):
        rp = torch.empty(N + 1, dtype=torch.int32, device=dev); ci = torch.empty(cnt, dtype=torch.int32, device=dev)
        va = torch.empty(cnt, dtype=torch.float64, device=dev)
        pkg.check(lib.ab_csr_export(hh, rp.data_ptr(), ci.data_ptr(), va.data_ptr()))
        return rp, ci, va

    a, b = export(h.value, nnz), export(g.value, G.query()[2])
    out["assembly_bit_exact_vs_generator"] = bool(all(torch.equal(x, y) for x, y in zip(a, b)))
    del a, b
    alg = 24.0 * T + 4.0 * (N + 1) + 12.0 * nnz
    out["assembly"] = {"ms": ms, "algorithmic_GB": alg / 1e9, "GBps": alg / ms / 1e6, "frac_of_peak": alg / ms / 1e6 / PEAK,
                       "Mtriplets_per_s": T / ms / 1e3}
    v2 = torch.rand(T, dtype=torch.float64, device=dev)
    ms_u = ev_time(lambda: pkg.check(lib.ab_csr_update_values(h.value, v2.data_ptr())), reps=3, warm=1)
    out["update_values"] = {"ms": ms_u, "GBps": (12.0 * T + 4.0 * nnz + 8.0 * nnz) / ms_u / 1e6}
    pkg.check(lib.ab_csr_update_values(h.value, vv.data_ptr()))
    x = torch.rand(N, dtype=torch.float64, device=dev) * 2 - 1
    y = torch.empty_like(x); t = C.c_double(0)
    pkg.check(lib.ab_bench_spmv(h.value, x.data_ptr(), y.data_ptr(), 5, 0, C.byref(t)))
    pkg.check(lib.ab_bench_spmv(h.value, x.data_ptr(), y.data_ptr(), 50, 0, C.byref(t)))
    bs = 12.0 * nnz + 4.0 * (N + 1) + 16.0 * N
    out["spmv"] = {"ms": t.value, "GBps": bs / t.value / 1e6, "frac_of_peak": bs / t.value / 1e6 / PEAK}
    dy = torch.rand(N, dtype=torch.float64, device=dev)
    dx = torch.empty_like(x)
    ms_gd = ev_time(lambda: pkg.check(lib.ab_spmm_grad_dense(h.value, dy.data_ptr(), 1, dx.data_ptr())), reps=10, warm=2)
    out["grad_dense"] = {"ms": ms_gd, "GBps": bs / ms_gd / 1e6, "frac_of_peak": bs / ms_gd / 1e6 / PEAK,
                         "symmetric_check": float((dx - (A @ dy)).abs().max())}
    dv = torch.empty(T, dtype=torch.float64, device=dev)
    ms_gv = ev_time(lambda: pkg.check(lib.ab_spmm_grad_values(h.value, dy.data_ptr(), x.data_ptr(), 1, dv.data_ptr())), reps=5, warm=1)
    bgv = 8.0 * nnz + 4.0 * (N + 1) + 16.0 * N + 12.0 * T      # per-triplet output + slot map
    chk = float((dv - dy[ii - 1] * x[jj - 1]).abs().max())
    out["grad_values"] = {"ms": ms_gv, "GBps": bgv / ms_gv / 1e6, "frac_of_peak": bgv / ms_gv / 1e6 / PEAK, "max_abs_err": chk}
    A.close(); G.close()
    return out


def c4(n=4096):
    out = {"config": f"C4: inverse conductivity, 2-D 5-pt variable coefficient {n}^2, forward CG + adjoint gradient w.r.t. all CSR values"}
    h_ = 1.0 / (n + 1)
    xs = torch.arange(0, n + 2, device=dev, dtype=torch.float64) * h_
    X, Y = torch.meshgrid(xs, xs, indexing="ij")
    kext = (1 + X ** 2 + Y ** 2).contiguous()               # strong_scaling_benchmark/benchmark.jl:6-8
    Xi, Yi = X[1:-1, 1:-1], Y[1:-1, 1:-1]
    rhs = (-2 * Xi * (1 - Xi) * (Xi ** 2 + Yi ** 2 + 1) + 2 * Xi * (-Xi * Yi * (1 - Yi) + Yi * (1 - Xi) * (1 - Yi))
           - 2 * Yi * (1 - Yi) * (Xi ** 2 + Yi ** 2 + 1) + 2 * Yi * (-Xi * Yi * (1 - Xi) + Xi * (1 - Xi) * (1 - Yi)))
    f = (-(2 * h_ * h_) * rhs).reshape(-1).contiguous()     # sign flipped with the matrix (SPD form)
    h = C.c_int64(0)
    ms_gen = ev_time(lambda: (lib.ab_csr_destroy(h.value) if h.value else 0, pkg.check(lib.ab_poisson2d_csr(n, kext.data_ptr(), -1, C.byref(h)))), reps=1)
    A = pkg.SparseTensor.from_handle(h.value)
    N, _, nnz, _ = A.query()
    out.update(N=N, nnz=nnz, generator_ms=ms_gen)
    u = torch.empty(N, dtype=torch.float64, device=dev)
    it, rr = C.c_int(0), C.c_double(0)
    t0 = time.perf_counter()
    pkg.check(lib.ab_solve(h.value, b"CG", f.data_ptr(), u.data_ptr(), 1e-10, 200000, C.byref(it), C.byref(rr)))
    torch.cuda.synchronize(); t_f = time.perf_counter() - t0
    res = float(((A @ u) - f).norm() / f.norm())
    ib = 12.0 * nnz + 4.0 * (N + 1) + 13.0 * 8.0 * N          # SURVEY §8d algorithmic bytes per PCG iteration
    out["forward"] = {"iters": it.value, "relres": rr.value, "true_relres": res, "seconds": t_f, "iters_per_s": it.value / t_f,
                      "GBps": ib * it.value / t_f / 1e9, "frac_of_peak": ib * it.value / t_f / 1e9 / PEAK}
    uex = (Xi * (1 - Xi) * Yi * (1 - Yi)).reshape(-1)
    out["forward"]["max_err_vs_analytic_u"] = float((u - uex).abs().max())
    # loss = sum(u^2): grad_u = 2u ; adjoint solve + per-entry gradient
    gu = (2 * u).contiguous()
    gvv = torch.empty(nnz, dtype=torch.float64, device=dev); gf = torch.empty(N, dtype=torch.float64, device=dev)
    t0 = time.perf_counter()
    pkg.check(lib.ab_solve_grad(h.value, b"CG", gu.data_ptr(), u.data_ptr(), gvv.data_ptr(), gf.data_ptr(), 1e-10, 200000, C.byref(it), C.byref(rr)))
    torch.cuda.synchronize(); t_b = time.perf_counter() - t0
    res_b = float(((A @ gf) - gu).norm() / gu.norm())
    out["adjoint"] = {"iters": it.value, "relres": rr.value, "true_relres": res_b, "seconds": t_b, "iters_per_s": it.value / t_b,
                      "frac_of_peak": ib * it.value / t_b / 1e9 / PEAK}
    # gradient w.r.t. kext through the generator adjoint (GetPoissonGrad): chain rule check by FD on one kappa entry
    uext = torch.zeros((n + 2, n + 2), dtype=torch.float64, device=dev); uext[1:-1, 1:-1] = u.view(n, n)
    gk = torch.empty((n + 2) * (n + 2), dtype=torch.float64, device=dev)
    xg = (-gf).contiguous()                                  # mpiops.jl:74: x = -(B \ du)
    ms_pg = ev_time(lambda: pkg.check(lib.ab_poisson2d_grad(n, xg.data_ptr(), uext.data_ptr(), -1, gk.data_ptr())), reps=5, warm=1)
    out["poisson_grad_ms"] = ms_pg
    out["backward_over_forward"] = t_b / t_f                 # reference: "typically less than twice" (mpi_benchmark.md:82)
    A.close()
    return out


def c5(rows=50_000_000, mean=16, k=32):
    out = {"config": f"C5: power-law rows={rows} mean nnz/row={mean}, SpMM k={k} + value gradient"}
    g = torch.Generator(device=dev).manual_seed(233)
    u = torch.rand(rows, device=dev, generator=g, dtype=torch.float64).clamp_min(1e-12)
    deg = torch.floor(u ** (-1.0 / 1.1)).clamp(1, 1e5)       # Zipf-like tail, alpha = 2.1
    deg = torch.clamp((deg * (mean / float(deg.mean()))).round(), 1, 1e5).to(torch.int64)
    del u
    rowptr = torch.zeros(rows + 1, dtype=torch.int64, device=dev)
    torch.cumsum(deg, 0, out=rowptr[1:])
    nnz = int(rowptr[-1].item())
    out.update(nnz=nnz, max_row=int(deg.max().item()), mean_row=nnz / rows)
    del deg
    rowptr = rowptr.to(torch.int32)
    colind = torch.randint(0, rows, (nnz,), device=dev, generator=g, dtype=torch.int32)
    values = torch.rand(nnz, device=dev, generator=g, dtype=torch.float64) * 2 - 1
    h = C.c_int64(0)
    pkg.check(lib.ab_csr_create(rows, rows, nnz, rowptr.data_ptr(), colind.data_ptr(), values.data_ptr(), C.byref(h)))
    X = torch.rand((rows, k), device=dev, generator=g, dtype=torch.float64) * 2 - 1
    Y = torch.empty((rows, k), device=dev, dtype=torch.float64)
    ms = ev_time(lambda: pkg.check(lib.ab_spmm(h.value, 0, X.data_ptr(), k, Y.data_ptr())), reps=3, warm=1)
    comp = 12.0 * nnz + 4.0 * (rows + 1) + 8.0 * k * 2 * rows
    gath = 12.0 * nnz + 4.0 * (rows + 1) + 8.0 * k * (nnz + rows)
    # property check on a sample of rows against torch gathers
    samp = torch.randint(0, rows, (2000,), device=dev, generator=g)
    err = 0.0
    rp64 = rowptr.to(torch.int64)
    for r in samp[:200].tolist():
        a, b = int(rp64[r]), int(rp64[r + 1])
        ref = (values[a:b, None] * X[colind[a:b].long()]).sum(0)
        err = max(err, float((Y[r] - ref).abs().max() / (ref.abs().max() + 1e-300)))
    out["spmm"] = {"ms": ms, "compulsory_GB": comp / 1e9, "GBps_compulsory": comp / ms / 1e6, "frac_compulsory": comp / ms / 1e6 / PEAK,
                   "no_reuse_gather_GB": gath / 1e9, "GBps_gather_model": gath / ms / 1e6, "frac_gather_model": gath / ms / 1e6 / PEAK,
                   "max_rel_err_sampled_rows": err}
    dv = torch.empty(nnz, device=dev, dtype=torch.float64)
    ms2 = ev_time(lambda: pkg.check(lib.ab_spmm_grad_values(h.value, Y.data_ptr(), X.data_ptr(), k, dv.data_ptr())), reps=2, warm=1)
    e = torch.randint(0, nnz, (1000,), device=dev, generator=g)
    rows_of = torch.searchsorted(rp64, e, right=True) - 1
    ref = (Y[rows_of] * X[colind[e].long()]).sum(1)
    gath2 = 4.0 * nnz + 4.0 * (rows + 1) + 8.0 * k * (nnz + rows) + 8.0 * nnz     # X row per entry, dY row once per row
    out["sddmm"] = {"ms": ms2, "GBps_compulsory": comp / ms2 / 1e6, "frac_compulsory": comp / ms2 / 1e6 / PEAK,
                    "no_reuse_gather_GB": gath2 / 1e9, "GBps_gather_model": gath2 / ms2 / 1e6,
                    "frac_gather_model": gath2 / ms2 / 1e6 / PEAK,
                    "max_rel_err_sampled": float(((dv[e] - ref).abs() / (ref.abs() + 1e-300)).max())}
    pkg.check(lib.ab_csr_destroy(h.value))
    return out


if __name__ == "__main__":
    path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/configs.json"
    which = (sys.argv[2] if len(sys.argv) > 2 else "c2,c4,c5").split(",")
    scale = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    res = {"peak_GBps": PEAK}
    for name in which:
        torch.cuda.empty_cache()
        t0 = time.time()
        try:
            if name == "c2":
                res[name] = c2(int(round(256 * scale)))
            elif name == "c4":
                res[name] = c4(int(round(4096 * scale)))
            elif name == "c5":
                res[name] = c5(int(50_000_000 * scale))
        except Exception as e:  # keep going: one section failing must not lose the others
            res[name] = {"error": repr(e)}
        res[name]["wall_s"] = time.time() - t0
        json.dump(res, open(path, "w"), indent=1)
        print(name, json.dumps(res[name])[:2000], flush=True)
