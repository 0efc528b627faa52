"""2-rank debug: which Krylov method does 'BoomerAMG' (auto) pick on the distributed stencil, and its iteration counts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
import __graft_entry__ as ge
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
pkg = ge.load_package(); pkg.mpi_init(rank, world, lr)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
pkg.check(pkg.lib.ab_set_option(b"verbose", 1.0))
A = pkg.mpi_SparseTensor.poisson3d(n, n, n, rank, world)
b = A @ torch.ones(A.nloc, dtype=torch.float64, device=dev)
for m in ("CG", "BoomerAMG", "GMRES", "CG"):
    u, it, rr = A.solve(b, m, tol=1e-11)
    if rank == 0: print(m, it, rr, flush=True)
pkg.mpi_finalize(); dist.destroy_process_group()
