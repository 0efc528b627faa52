"""Multi-GPU parity check, launched under torchrun on a real multi-GPU box (not collected by pytest; the
pytest form is tests/test_gpu_dist.py, the driver-visible form is the "parity" block of bench.py at N > 1):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \\
        --master-port 29533 tests/dist_check.py [grid=96]

Runs adcme.jl_b200/selfcheck.py:dist_parity on every rank and prints every check with its value."""
import json
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
    from adcme_jl_b200 import selfcheck

    res = selfcheck.dist_parity(pkg, rank, world, dev, n=n)
    for r in range(world):
        if r == rank:
            for k, v in res["checks"].items():
                print(f"rank {rank}/{world}: {k} = {v:.3e} {'FAIL' if k in res['failed'] else 'OK'}", flush=True)
        dist.barrier()
    if rank == 0:
        print(json.dumps({"ok": res["ok"], "world": world, "grid": n, "n_checks": len(res["checks"])}), flush=True)
    pkg.mpi_finalize()
    dist.destroy_process_group()
    sys.exit(0 if res["ok"] else 1)


if __name__ == "__main__":
    main()
