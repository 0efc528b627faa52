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
