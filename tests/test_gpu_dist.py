"""Multi-rank parity under pytest: two ranks are spawned (one per GPU) when the box has at least two devices;
on a single-GPU box the test is skipped and the same checks run inside `bench.py` at N > 1 (its JSON line
carries them as "parity") and through tests/dist_check.py under torchrun.

Covers what a single process cannot: the peer-memory halo push + in-kernel halo wait of both one-launch SpMV
kernels (row-pattern and x-window), the hand-written 2-value all-reduce and its folded-in forms, the NCCL
fallback, distributed BiCGSTAB on an unsymmetric matrix, the distributed transpose, the COO export and the
adjoint solve through A' — each against scipy / the single-GPU solver / a matrix-free stencil."""
import json
import os
import socket
import sys
import tempfile

import pytest

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, outdir, grid):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, REPO)
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    pkg = ge.load_package()
    pkg.mpi_init(rank, world, rank)
    from adcme_jl_b200 import selfcheck

    res = selfcheck.dist_parity(pkg, rank, world, dev, n=grid)
    json.dump(res, open(os.path.join(outdir, f"rank{rank}.json"), "w"))
    dist.barrier()
    pkg.mpi_finalize()
    dist.destroy_process_group()
    if not res["ok"]:
        raise SystemExit(1)


@pytest.mark.timeout(600)
def test_two_rank_parity(pkg):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (the same checks run in bench.py at N > 1 and in tests/dist_check.py)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as td:
        try:
            mp.spawn(_worker, args=(2, port, td, 64), nprocs=2, join=True)
            ok = True
        except Exception as e:          # a rank exited non-zero: show which checks failed
            ok = False
            err = repr(e)
        res = [json.load(open(os.path.join(td, f))) for f in sorted(os.listdir(td))]
        assert ok, f"{err}\n" + json.dumps([r.get("failed") for r in res])
        for r in res:
            assert r["ok"] and not r["failed"], r["failed"]
            assert len(r["checks"]) >= 30


def test_single_rank_runs_the_same_checks(pkg):
    """world = 1 through the very same code: distributed BiCGSTAB, symmetry check, transpose (self-exchange), COO export,
    true-residual vetting — so the driver's one-GPU test run exercises them too."""
    import torch

    from adcme_jl_b200 import selfcheck

    pkg.mpi_init(0, 1, torch.cuda.current_device())
    res = selfcheck.dist_parity(pkg, 0, 1, torch.device("cuda", torch.cuda.current_device()), n=40)
    assert res["ok"], res["failed"]
    assert len(res["checks"]) >= 20


def test_dist_solve_reports_failure_instead_of_a_wrong_answer(pkg):
    """"CG" on a symmetric INDEFINITE system must not return AB_OK with a wrong answer (true-residual vetting),
    while "BoomerAMG"/auto falls back to BiCGSTAB and solves it."""
    import numpy as np
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    import torch

    pkg.mpi_init(0, 1, torch.cuda.current_device())
    M = 400
    rng = np.random.default_rng(5)
    d = np.where(np.arange(M) % 2 == 0, 4.0, -4.0)                 # indefinite diagonal, still diagonally dominant
    G = (sp.diags(d) + sp.diags(np.full(M - 1, 1.0), 1) + sp.diags(np.full(M - 1, 1.0), -1)).tocsr()
    coo = G.tocoo()
    L = pkg.SparseTensor(coo.row.astype(np.int64) + 1, coo.col.astype(np.int64) + 1, coo.data, M, M)
    D = pkg.mpi_SparseTensor(L, 0, M - 1, M)
    f = rng.uniform(size=M)
    u0 = spla.spsolve(G.tocsc(), f)
    u, it, rr = D.solve(f, "BoomerAMG", tol=1e-12)
    assert np.max(np.abs(u - u0)) / np.max(np.abs(u0)) <= 1e-9 and rr <= 1e-11
    try:
        u, it, rr = D.solve(f, "CG", tol=1e-12, maxit=500)
        assert np.max(np.abs(u - u0)) / np.max(np.abs(u0)) <= 1e-8          # if CG happened to get there, it must be right
    except pkg.ADCMEError as e:
        assert e.code in (-5, -6)
    D.close()
