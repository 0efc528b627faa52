"""Multi-GPU parity check, launched under torchrun on a real multi-GPU box (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/dist_check.py [grid=96]

Each rank owns a z-slab of the n^3 7-point Poisson matrix.  Checks, against quantities every rank can
compute alone: (1) distributed SpMV == matrix-free stencil on the same seeded global vector,
(2) distributed Jacobi-PCG solve of A u = A*1 returns ones with relres <= 1e-10, and (3) the iterate
after a fixed number of iterations equals the single-GPU solver's iterate to 1e-12."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import __graft_entry__ as ge  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    pkg = ge.load_package()
    pkg.mpi_init(rank, world, lr)
    N = n ** 3
    A = pkg.mpi_SparseTensor.poisson3d(n, n, n, rank, world)
    lo, hi = A.ilower, A.iupper
    g = torch.Generator(device=dev).manual_seed(233)
    x = torch.rand(N, dtype=torch.float64, device=dev, generator=g) * 2 - 1      # same on every rank
    X = x.view(n, n, n)
    S = 6.0 * X
    S[1:] -= X[:-1]; S[:-1] -= X[1:]
    S[:, 1:] -= X[:, :-1]; S[:, :-1] -= X[:, 1:]
    S[:, :, 1:] -= X[:, :, :-1]; S[:, :, :-1] -= X[:, :, 1:]
    y = A @ x[lo:hi + 1].contiguous()
    e1 = float((y - S.reshape(-1)[lo:hi + 1]).abs().max() / S.abs().max())
    b = A @ torch.ones(A.nloc, dtype=torch.float64, device=dev)
    u, it, rr = A.solve(b, tol=1e-11)
    e2 = float((u - 1).abs().max())
    # single-GPU reference iterate on the full system (every rank can afford it at test sizes)
    h = C.c_int64(0)
    pkg.check(pkg.lib.ab_poisson3d_csr(n, n, n, 0, N, C.byref(h)))
    Af = pkg.SparseTensor.from_handle(h.value)      # keep the wrapper alive: its finaliser frees the handle
    bf = Af @ torch.ones(N, dtype=torch.float64, device=dev)
    uf = torch.empty_like(bf)
    r = C.c_double(0)
    pkg.check(pkg.lib.ab_cg_fixed(h.value, bf.data_ptr(), uf.data_ptr(), 40, C.byref(r)))
    ud, rd = A.cg_fixed(b, 40)
    e3 = float((ud - uf[lo:hi + 1]).abs().max() / uf.abs().max())
    ok = e1 <= 1e-12 and rr <= 1e-11 and e2 <= 1e-7 and e3 <= 1e-12 and abs(rd - r.value) <= 1e-10 * r.value
    print(f"rank {rank}/{world}: rows [{lo},{hi}] spmv_err {e1:.2e} solve iters {it} relres {rr:.2e} err {e2:.2e} "
          f"fixed-iterate err {e3:.2e} {'OK' if ok else 'FAIL'}", flush=True)
    # (4) a GENERAL matrix (irregular halos, no stencil structure): random symmetric diagonally dominant
    # band matrix, rows split unevenly; the distributed solve must match scipy's direct solve
    # (the reference's own check is a 5x5 A*A' system through Hypre: test_mpi_tensor_solve.jl:5-26)
    import numpy as np
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    M = 3001
    rng = np.random.default_rng(233)
    rows = np.repeat(np.arange(M), 6)
    cols = np.clip(rows + rng.integers(-400, 401, rows.size), 0, M - 1)
    B = sp.coo_matrix((rng.uniform(-1, 1, rows.size), (rows, cols)), shape=(M, M)).tocsr()
    G = (B + B.T).tocsr()
    G = (G + sp.diags(np.asarray(abs(G).sum(axis=1)).ravel() + 1.0)).tocsr()
    G.sort_indices()
    glo, ghi = pkg.row_partition(M, world, rank)
    loc = G[glo:ghi + 1].tocoo()
    L = pkg.SparseTensor(loc.row.astype(np.int64) + 1, loc.col.astype(np.int64) + 1, loc.data, ghi - glo + 1, M)
    D = pkg.mpi_SparseTensor(L, glo, ghi, M)
    xg = rng.uniform(-1, 1, M)
    yg = D @ xg[glo:ghi + 1]
    e4 = float(np.max(np.abs(yg - (G @ xg)[glo:ghi + 1])) / np.max(np.abs(G @ xg)))
    fg = rng.uniform(size=M)
    ug, itg, rrg = D.solve(fg[glo:ghi + 1], tol=1e-12)
    u0 = spla.spsolve(G.tocsc(), fg)
    e5 = float(np.max(np.abs(ug - u0[glo:ghi + 1])) / np.max(np.abs(u0)))
    # adjoint: x = G^-1 g (G symmetric), grad_f = x, grad_vals[e] = -x[row_e] u[col_e] over this rank's CSR entries
    gg = rng.uniform(-1, 1, M)
    gvals, gfl = D.solve_grad(gg[glo:ghi + 1], u0[glo:ghi + 1], tol=1e-13)
    x0 = spla.spsolve(G.tocsc(), gg)
    Gl = G[glo:ghi + 1].tocsr(); Gl.sort_indices()
    rows_g = np.repeat(np.arange(glo, ghi + 1), np.diff(Gl.indptr))
    gv0 = -x0[rows_g] * u0[Gl.indices]
    e6 = float(np.max(np.abs(gfl - x0[glo:ghi + 1])) / np.max(np.abs(x0)))
    e7 = float(np.max(np.abs(gvals - gv0)) / np.max(np.abs(gv0)))
    ok4 = e4 <= 1e-13 and e5 <= 1e-9 and rrg <= 1e-12 and e6 <= 1e-9 and e7 <= 1e-9
    print(f"rank {rank}/{world}: general matrix rows [{glo},{ghi}] spmv_err {e4:.2e} solve iters {itg} relres {rrg:.2e} "
          f"err vs direct {e5:.2e} adjoint grad_f err {e6:.2e} grad_vals err {e7:.2e} {'OK' if ok4 else 'FAIL'}", flush=True)
    ok = ok and ok4
    D.close()
    t = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(t)
    A.close()
    pkg.mpi_finalize()
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 0 else 1)


if __name__ == "__main__":
    main()
