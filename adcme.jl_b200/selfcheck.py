"""Multi-rank parity checks of the row-block distributed path (mpi_SparseTensor mirror), callable from
bench.py (world > 1, before timing), from the `-m gpu` tests (two spawned ranks) and from tests/dist_check.py.

Every rank calls `dist_parity` (collective).  References are quantities each rank can compute alone:
a matrix-free torch stencil, the single-GPU solver of this library on the full system (fixed-iterate
comparison), and scipy's direct solver on small general matrices — never the oracle, never a CPU fallback
of the product.  Returns {"ok": bool, "checks": {name: value}, "failed": [names]} identical on all ranks.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def _set(lib, check, **kv):
    for k, v in kv.items():
        check(lib.ab_set_option(k.encode(), float(v)))


def dist_parity(pkg, rank: int, world: int, dev, n: int = 96, variants: bool = True, general: bool = True):
    import torch
    import torch.distributed as dist

    lib, check = pkg.lib, pkg.check
    out, failed = {}, []

    def record(name, value, ok):
        out[name] = float(value)
        if not ok:
            failed.append(name)

    N = n ** 3
    # ---- (1) stencil SpMV vs a matrix-free stencil; (2) full solve; (3) fixed iterate vs the single-GPU solver
    g = torch.Generator(device=dev).manual_seed(233)
    x = torch.rand(N, dtype=torch.float64, device=dev, generator=g) * 2 - 1      # same on every rank
    X = x.view(n, n, n)
    S = 6.0 * X
    S[1:] -= X[:-1]; S[:-1] -= X[1:]
    S[:, 1:] -= X[:, :-1]; S[:, :-1] -= X[:, 1:]
    S[:, :, 1:] -= X[:, :, :-1]; S[:, :, :-1] -= X[:, :, 1:]
    h = C.c_int64(0)
    check(lib.ab_poisson3d_csr(n, n, n, 0, N, C.byref(h)))
    Af = pkg.SparseTensor.from_handle(h.value)      # keep the wrapper alive: its finaliser frees the handle
    bf = Af @ torch.ones(N, dtype=torch.float64, device=dev)
    uf = torch.empty_like(bf)
    r1 = C.c_double(0)
    check(lib.ab_cg_fixed(h.value, bf.data_ptr(), uf.data_ptr(), 40, C.byref(r1)))

    def stencil_case(tag, solve: bool):
        A = pkg.mpi_SparseTensor.poisson3d(n, n, n, rank, world)
        lo, hi = A.ilower, A.iupper
        y = A @ x[lo:hi + 1].contiguous()
        e1 = float((y - S.reshape(-1)[lo:hi + 1]).abs().max() / S.abs().max())
        record(f"{tag}spmv_vs_matrix_free", e1, e1 <= 1e-12)
        b = A @ torch.ones(A.nloc, dtype=torch.float64, device=dev)
        ud, rd = A.cg_fixed(b, 40)
        e3 = float((ud - uf[lo:hi + 1]).abs().max() / uf.abs().max())
        record(f"{tag}fixed_iterate_vs_single_gpu", e3, e3 <= 1e-12 and abs(rd - r1.value) <= 1e-10 * r1.value)
        if solve:
            u, it, rr = A.solve(b, "CG", tol=1e-11)
            e2 = float((u - 1).abs().max())
            record(f"{tag}solve_true_relres", rr, rr <= 1e-10)
            record(f"{tag}solve_err_vs_ones", e2, e2 <= 1e-7)
            u2, it2, rr2 = A.solve(b, "BoomerAMG", tol=1e-11)       # auto: symmetric across ranks -> CG
            record(f"{tag}auto_is_cg", abs(it2 - it), abs(it2 - it) <= max(3, it // 20) and rr2 <= 1e-10)
        A.close()

    stencil_case("", True)
    if variants and world > 1:
        # every combination of data path the solver can take: NCCL vs peer memory, separate halo-wait launch vs the
        # in-kernel wait (row-pattern kernel and x-window kernel), all-reduces as launches vs folded into K2'/K3'
        # ... and K3 pushing the next halos while it sweeps (k3h) vs the separate halo_push_kernel launch
        combos = [dict(dist_p2p=0, dist_fused_wait=0, dist_fused_reduce=0, dist_fused_push=0, spmv_rowwin=1, dist_k3_halo=1, dist_ll=1),
                  dict(dist_p2p=1, dist_fused_wait=0, dist_fused_reduce=0, dist_fused_push=0, spmv_rowwin=1, dist_k3_halo=0, dist_ll=0),
                  dict(dist_p2p=1, dist_fused_wait=0, dist_fused_reduce=0, dist_fused_push=0, spmv_rowwin=1, dist_k3_halo=1, dist_ll=1),
                  dict(dist_p2p=1, dist_fused_wait=1, dist_fused_reduce=0, dist_fused_push=0, spmv_rowwin=0, dist_k3_halo=1, dist_ll=1),
                  dict(dist_p2p=1, dist_fused_wait=1, dist_fused_reduce=1, dist_fused_push=0, spmv_rowwin=0, dist_k3_halo=1, dist_ll=1),
                  dict(dist_p2p=1, dist_fused_wait=1, dist_fused_reduce=1, dist_fused_push=1, spmv_rowwin=0, dist_k3_halo=1, dist_ll=1),
                  dict(dist_p2p=1, dist_fused_wait=1, dist_fused_reduce=0, dist_fused_push=0, spmv_rowwin=1, dist_k3_halo=0, dist_ll=1),
                  dict(dist_p2p=1, dist_fused_wait=1, dist_fused_reduce=0, dist_fused_push=0, spmv_rowwin=1, dist_k3_halo=1, dist_ll=1),
                  dict(dist_p2p=1, dist_fused_wait=1, dist_fused_reduce=1, dist_fused_push=0, spmv_rowwin=1, dist_k3_halo=1, dist_ll=1),
                  dict(dist_p2p=1, dist_fused_wait=1, dist_fused_reduce=1, dist_fused_push=1, spmv_rowwin=1, dist_k3_halo=1, dist_ll=1)]
        saved = {}
        for k in combos[0]:
            v = C.c_double(0)
            saved[k] = v.value if lib.ab_get_option(k.encode(), C.byref(v)) == 0 else None
        defaults = dict(dist_p2p=1, dist_fused_wait=1, dist_fused_reduce=0, dist_fused_push=0, spmv_rowwin=1, dist_k3_halo=1, dist_ll=1)
        defaults.update({k: v for k, v in saved.items() if v is not None})       # A/B switches of the caller survive
        for cmb in combos:
            _set(lib, check, **cmb)
            tag = "p2p%d_wait%d_red%d_push%d_win%d_k3h%d_ll%d:" % (cmb["dist_p2p"], cmb["dist_fused_wait"], cmb["dist_fused_reduce"],
                                                                  cmb["dist_fused_push"], cmb["spmv_rowwin"], cmb["dist_k3_halo"], cmb["dist_ll"])
            try:
                stencil_case(tag, False)
            finally:
                _set(lib, check, **defaults)
    Af.close()
    if general:
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla

        # ---- (4) a GENERAL matrix (irregular halos, no stencil structure, uneven row split): symmetric diagonally
        # dominant band matrix; solve and adjoint vs scipy's direct solve (the reference's own check is a 5x5
        # A*A' system through Hypre: test_mpi_tensor_solve.jl:5-26)
        M = 3001
        rng = np.random.default_rng(233)
        rows = np.repeat(np.arange(M), 6)
        cols = np.clip(rows + rng.integers(-400, 401, rows.size), 0, M - 1)
        B = sp.coo_matrix((rng.uniform(-1, 1, rows.size), (rows, cols)), shape=(M, M)).tocsr()
        Gs = (B + B.T).tocsr()
        Gs = (Gs + sp.diags(np.asarray(abs(Gs).sum(axis=1)).ravel() + 1.0)).tocsr()
        # ---- (5) the same with an unsymmetric part: BiCGSTAB ("GMRES"), the distributed transpose for the adjoint
        U = (sp.triu(B, 1) * 0.5).tocsr()
        Gu = (Gs + U + sp.diags(np.asarray(abs(U).sum(axis=1)).ravel())).tocsr()      # still strictly diagonally dominant
        glo, ghi = pkg.row_partition(M, world, rank)
        for tag, G, methods in (("sym:", Gs, ("CG", "BoomerAMG", "GMRES")), ("unsym:", Gu, ("GMRES", "BoomerAMG"))):
            G.sort_indices()
            loc = G[glo:ghi + 1].tocoo()
            L = pkg.SparseTensor(loc.row.astype(np.int64) + 1, loc.col.astype(np.int64) + 1, loc.data, ghi - glo + 1, M)
            D = pkg.mpi_SparseTensor(L, glo, ghi, M)
            xg = rng.uniform(-1, 1, M)
            yg = D @ xg[glo:ghi + 1]
            e4 = float(np.max(np.abs(yg - (G @ xg)[glo:ghi + 1])) / np.max(np.abs(G @ xg)))
            record(tag + "spmv_vs_scipy", e4, e4 <= 1e-13)
            fg = rng.uniform(size=M)
            u0 = spla.spsolve(G.tocsc(), fg)
            for m in methods:
                ug, itg, rrg = D.solve(fg[glo:ghi + 1], m, tol=1e-12)
                e5 = float(np.max(np.abs(ug - u0[glo:ghi + 1])) / np.max(np.abs(u0)))
                record(f"{tag}solve[{m}]_vs_direct", e5, e5 <= 1e-9 and rrg <= 1e-11)
            # adjoint: x = G^-T g, grad_f = x, grad_vals[e] = -x[row_e] u[col_e] over this rank's CSR entries
            gg = rng.uniform(-1, 1, M)
            gvals, gfl = D.solve_grad(gg[glo:ghi + 1], u0[glo:ghi + 1], tol=1e-13)
            x0 = spla.spsolve(G.T.tocsc(), gg)
            Gl = G[glo:ghi + 1].tocsr(); Gl.sort_indices()
            rows_g = np.repeat(np.arange(glo, ghi + 1), np.diff(Gl.indptr))
            gv0 = -x0[rows_g] * u0[Gl.indices]
            e6 = float(np.max(np.abs(gfl - x0[glo:ghi + 1])) / np.max(np.abs(x0)))
            e7 = float(np.max(np.abs(gvals - gv0)) / np.max(np.abs(gv0)))
            record(tag + "adjoint_grad_f", e6, e6 <= 1e-9)
            record(tag + "adjoint_grad_vals", e7, e7 <= 1e-9)
            # SparseTensor(sp): the local block back as COO (MPIGetMatrix), bit-exact
            idx2, vals = D.coo()
            glc = Gl.tocoo()
            same = (np.array_equal(idx2[:, 0], glc.row) and np.array_equal(idx2[:, 1], glc.col) and np.array_equal(vals, glc.data))
            record(tag + "export_coo_exact", 0.0 if same else 1.0, same)
            # adjoint(A): this rank's rows of G', bit-exact (values only move)
            Dt = D.adjoint()
            it2, vt = Dt.coo()
            Gt = G.T.tocsr()[glo:ghi + 1].tocsr(); Gt.sort_indices(); gtc = Gt.tocoo()
            same_t = (np.array_equal(it2[:, 0], gtc.row) and np.array_equal(it2[:, 1], gtc.col) and np.array_equal(vt, gtc.data))
            record(tag + "transpose_exact", 0.0 if same_t else 1.0, same_t)
            yt = Dt @ xg[glo:ghi + 1]
            e8 = float(np.max(np.abs(yt - (G.T @ xg)[glo:ghi + 1])) / np.max(np.abs(G.T @ xg)))
            record(tag + "transpose_spmv", e8, e8 <= 1e-13)
            Dt.close()
            D.close()
    # agree on the verdict
    t = torch.tensor([len(failed)], device=dev, dtype=torch.int64)
    if world > 1:
        dist.all_reduce(t)
    return {"ok": int(t.item()) == 0, "checks": out, "failed": failed, "grid": n, "world": world}
