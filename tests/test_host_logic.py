"""Host-side logic of the N>1 path, exercised with world_size-2 gloo process groups on CPU:
row partition (Hypre IJ layout), the NCCL-id broadcast plumbing of mpi_init, and bench.py's
work sharding + max-over-ranks reduction."""
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_partition_tiles_the_rows(pkg):
    for N, world, align in ((512 ** 3, 8, 512 * 512), (4096 ** 2, 4, 4096), (1000, 3, 1), (64, 8, 16), (10, 1, 1)):
        parts = [pkg.row_partition(N, world, r, align) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == N - 1
        for (a, b), (c, d) in zip(parts, parts[1:]):
            assert c == b + 1
        for a, b in parts[:-1]:
            assert a % align == 0 and (b + 1) % align == 0
    # 512^3 over 8 ranks: 64 planes each
    assert pkg.row_partition(512 ** 3, 8, 3, 512 * 512) == (3 * 64 * 512 * 512, 4 * 64 * 512 * 512 - 1)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, REPO)
    import torch
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bench

        # 1. the id broadcast used by mpi_init: rank 0's 128 bytes reach everyone
        t = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            t = torch.arange(128, dtype=torch.uint8)
        dist.broadcast(t, src=0)
        ok_id = bool((t == torch.arange(128, dtype=torch.uint8)).all())
        # 2. bench sharding: ranks own disjoint z-slabs that tile the grid
        lo, hi = bench.slab_rows(16, world, rank)
        # 3. max-over-ranks timing reduction
        mx = bench.max_over_ranks(float(10 * (rank + 1)), device="cpu")
        tot = bench.sum_over_ranks(float(hi - lo + 1), device="cpu")
        q.put((rank, ok_id, lo, hi, mx, tot))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_plumbing():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, ok0, lo0, hi0, mx0, tot0), (r1, ok1, lo1, hi1, mx1, tot1) = res
    assert ok0 and ok1
    assert lo0 == 0 and lo1 == hi0 + 1 and hi1 == 16 ** 3 - 1
    assert mx0 == mx1 == 20.0
    assert tot0 == tot1 == 16 ** 3


# ---------------------------------------------------------------------------------------------
# host mirror logic that needs no device: Julia-style index arguments, spdiag / spzero triplets,
# Jacobian-handle reuse decision, option defaults (src/sparse.jl:303-314,540-575,631-658; options.jl)
def test_julia_style_index_arguments(pkg):
    import importlib

    st = importlib.import_module(pkg.__name__ + ".structure")
    il = st._index_list
    assert il(slice(None), 5)[0].tolist() == [1, 2, 3, 4, 5] and il(None, 3)[0].tolist() == [1, 2, 3]
    assert il(slice(2, 3), 10)[0].tolist() == [2, 3]                  # 2:3 is inclusive, 1-based
    assert il(slice(1, 9, 4), 10)[0].tolist() == [1, 5, 9]            # 1:4:9
    assert il(7, 10) == (pytest.approx(np.array([7])), True)          # integer index squeezes the dimension
    assert il(np.array([False, True, True]), 3)[0].tolist() == [2, 3]  # findall(mask)
    assert il([4, 1], 5)[0].tolist() == [4, 1] and il(range(2, 5), 9)[0].tolist() == [2, 3, 4]


def test_spdiag_spzero_triplets_without_device(pkg):
    A = pkg.spdiag(4)
    assert (A.m, A.n) == (4, 4) and np.asarray(A.ii).tolist() == [1, 2, 3, 4] and np.asarray(A.vv).tolist() == [1.0] * 4
    D = pkg.spdiag(np.array([3.0, 5.0]))
    assert D._diag and np.asarray(D.jj).tolist() == [1, 2]
    B = pkg.spdiag(5, (0, np.arange(5.0)), (-2, np.ones(3)), (1, 2 * np.ones(4)))     # sparse.jl:631-658
    dense = np.zeros((5, 5))
    np.add.at(dense, (np.asarray(B.ii) - 1, np.asarray(B.jj) - 1), np.asarray(B.vv))
    assert np.array_equal(dense, np.diag(np.arange(5.0)) + np.diag(np.ones(3), -2) + np.diag(2 * np.ones(4), 1))
    with pytest.raises(ValueError):
        pkg.spdiag(5, (0, np.ones(4)))
    Z = pkg.spzero(3, 7)
    assert Z.shape == (3, 7) and len(Z.vv) == 0
    with pytest.raises(ValueError):
        pkg.hcat(pkg.spzero(3, 2), pkg.spzero(4, 2))                  # shape check happens before any device call
    with pytest.raises(ValueError):
        pkg.spzero(3, 3) + pkg.spzero(3, 4)


def test_newton_options_and_jacobian_reuse_decision(pkg):
    import importlib

    op = importlib.import_module(pkg.__name__ + ".optim")
    pkg.reset_default_options()
    nr = pkg.options.newton_raphson                                    # options.jl:20-44
    assert (nr.max_iter, nr.rtol, nr.tol, nr.LM, nr.linesearch) == (100, 1e-12, 1e-12, 0.0, False)
    ls = nr.linesearch_options
    assert (ls.c1, ls.rho_hi, ls.rho_lo, ls.iterations, ls.alpha_initial) == (1e-4, 0.5, 0.1, 1000, 1.0)
    i, j = np.array([1, 2]), np.array([1, 2])
    assert op._same_indices(i, i) and op._same_indices(i, i.copy()) and not op._same_indices(i, np.array([1, 3]))
    assert not op._same_indices(i, np.array([1, 2, 3])) and not op._same_indices(i, None)
    with pytest.raises(ValueError):
        pkg.newton_raphson(lambda th, u: (u, None), np.ones((2, 2)))  # "Initial guess must be a vector"


def test_bench_affinity_helper_is_a_noop_without_a_gpu():
    """bench.py pins its e2e host buffers next to the rank's GPU; without NVML / a GPU the context manager must do nothing
    and must always restore the affinity it found."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    before = os.sched_getaffinity(0)
    with bench.near_gpu_affinity(0):
        inside = os.sched_getaffinity(0)
        assert inside <= before and len(inside) > 0
    assert os.sched_getaffinity(0) == before


def _zmarch_plan(pkg, M, m, grid, excl=None, fits_g_max=4, fits_l2=0, g_opt=0, len_opt=0, bal_opt=-1):
    import ctypes as C

    cap = 40000
    out3 = (C.c_int * 3)()
    first = np.zeros(cap, np.uint32); count = np.zeros(cap, np.int32); rows = np.zeros(cap, np.int32)
    cta = np.zeros(cap, np.int32); rest = np.zeros(cap, np.int32)
    ni, nc, nr = C.c_int64(0), C.c_int64(0), C.c_int64(0)
    ex = None if excl is None else np.ascontiguousarray(excl, np.uint8)
    pkg.check(pkg.lib.ab_test_zmarch_plan(M, m, grid, None if ex is None else ex.ctypes.data, 0 if ex is None else len(ex), fits_g_max,
                                          fits_l2, g_opt, len_opt, bal_opt, out3, cap, first.ctypes.data, count.ctypes.data,
                                          rows.ctypes.data, cta.ctypes.data, rest.ctypes.data, C.byref(ni), C.byref(nc), C.byref(nr)))
    n = ni.value
    return dict(G=out3[0], balanced=out3[1], seg=out3[2], first=first[:n].astype(np.int64), count=count[:n].astype(np.int64),
                rows=rows[:n].astype(np.int64), cta=cta[:nc.value + 1] if nc.value else None, rest=rest[:nr.value])


@pytest.mark.parametrize("M,planes,grid,kw", [
    (512 * 512, 512, 148, {}),                          # the headline grid: G = 4, round-robin z-segments
    (512 * 512, 64, 148, dict(fits_l2=1)),              # one rank's slab at N = 8: balanced partition
    (400 * 400, 400, 148, {}),                          # plane stride not a multiple of the tile: narrower last strip
    (100 * 100, 100, 148, {}),
    (1040, 1040, 148, {}),                              # 2-D grid with long lines: one strip per line
    (64 * 64, 64, 148, {}),
    (512 * 512, 512, 148, dict(g_opt=2)),
    (512 * 512, 64, 132, dict(bal_opt=1)),
])
def test_zmarch_plan_covers_every_row_exactly_once(pkg, M, planes, grid, kw):
    """Host-side planning of the marching SpMV (csrc/zmarch_plan.h through the host-only hook): the work items are chains of
    positions `M` rows apart; together they must cover every row of the matrix exactly once, no position wider than G tiles,
    and a balanced partition hands every CTA the same number of positions (+-1)."""
    m = M * planes
    p = _zmarch_plan(pkg, M, m, grid, **kw)
    assert p["G"] in (1, 2, 4) and len(p["first"]) > 0
    TR = 1024 * p["G"]
    seen = np.zeros(m + 1, np.int64)                    # difference array over rows
    for f, c, r in zip(p["first"], p["count"], p["rows"]):
        assert 1 <= r <= TR and c >= 1
        assert (f % M) % TR == 0 and r == min(TR, M - f % M)            # a strip of the plane, full or the narrow last one
        for q in range(c):
            a = f + q * M
            b = min(a + r, m)
            assert a < m
            seen[a] += 1; seen[b] -= 1
    cover = np.cumsum(seen[:-1])
    assert cover.min() == 1 and cover.max() == 1
    assert len(p["rest"]) == 0
    if p["balanced"]:
        cta = p["cta"]
        assert cta is not None and cta[0] == 0 and cta[-1] == len(p["first"]) and len(cta) - 1 <= grid
        per = [int(p["count"][cta[b]:cta[b + 1]].sum()) for b in range(len(cta) - 1)]
        assert max(per) - min(per) <= 1 and sum(per) == int(p["count"].sum())
    else:
        assert p["cta"] is None and p["seg"] >= 1 and p["count"].max() <= p["seg"]


def test_zmarch_plan_with_excluded_tiles_leaves_them_to_the_rest_list(pkg):
    """A distributed block: the tiles of the first and last plane read halo columns and are excluded; every other row is
    covered exactly once, the excluded tiles (and nothing else) form the rest list, chains are cut around them."""
    M, planes, grid = 512 * 512, 64, 148
    m = M * planes
    ntiles = m // 1024
    excl = np.zeros(ntiles, np.uint8)
    excl[: M // 1024] = 1
    excl[-(M // 1024):] = 1
    for kw in (dict(fits_l2=1), dict(bal_opt=0)):
        p = _zmarch_plan(pkg, M, m, grid, excl=excl, **kw)
        TR = 1024 * p["G"]
        assert M % TR == 0                               # strips coincide with tiles when tiles are excluded
        tiles = np.zeros(ntiles, np.int64)
        for f, c, r in zip(p["first"], p["count"], p["rows"]):
            for q in range(c):
                a = f + q * M
                tiles[a // 1024:(a + r) // 1024] += 1
        assert np.array_equal(tiles, 1 - excl.astype(np.int64))
        assert np.array_equal(np.sort(p["rest"]), np.nonzero(excl)[0])


def test_zmarch_plan_needs_tile_aligned_strips_when_tiles_are_excluded(pkg):
    M, planes = 400 * 400, 8                            # 160000 rows per plane: no multiple of 1024
    m = M * planes
    excl = np.zeros((m + 1023) // 1024, np.uint8)
    excl[0] = 1
    p = _zmarch_plan(pkg, M, m, 148, excl=excl)
    assert p["G"] == 0 and len(p["first"]) == 0          # the kernel steps aside (x-window kernel takes the matrix)


def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    """The driver launches `bench.py --impl reference` exactly like the GPU arm (torchrun for N > 1).  Rank 0 alone runs the CPU
    port and prints ONE JSON line; the other rank exits 0 without output; the thread count is set explicitly (torchrun exports
    OMP_NUM_THREADS=1 — the round-1 scaling run measured one core at N >= 2 because of it)."""
    import json
    import subprocess

    port = 31000 + (os.getpid() % 2000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "0", "--grid", "48"]
    out = subprocess.run(cmd, cwd=REPO, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["gpu_launches"] == 0
    assert d["unit"] == "CG iterations/s" and d["higher_is_better"] is True and d["value"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["single_thread_value"] > 0
    ncpu = len(os.sched_getaffinity(0))
    assert cb["cores"] == ncpu or ncpu == 1, (cb, ncpu)              # every host thread, whatever torchrun put into the environment
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
