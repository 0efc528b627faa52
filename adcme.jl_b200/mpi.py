"""Row-block distributed sparse path: mirror of mpi_SparseTensor and its `\\`
(/root/reference/src/mpi.jl:420-508) on one process per GPU.

torch.distributed is plumbing only (it broadcasts the 128-byte NCCL id that bootstraps the
library's own communicator); halos and inner products move inside libadcme_b200 (dist.cu).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from ._lib import as_array, check, empty_like_kind, lib, ptr
from .options import options
from .sparse import SparseTensor

_initialized = False
_rank, _world = 0, 1


def mpi_init(rank: int | None = None, world: int | None = None, device: int | None = None) -> None:
    """mpi_init (mpi.jl:16-31).  Under torchrun reads RANK / WORLD_SIZE / LOCAL_RANK."""
    global _initialized, _rank, _world
    rank = int(os.environ.get("RANK", 0)) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", 1)) if world is None else world
    device = int(os.environ.get("LOCAL_RANK", rank)) if device is None else device
    check(lib.ab_init(device))
    idbuf = (C.c_ubyte * 128)()
    if world > 1:
        import torch
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("mpi_init(world>1) needs torch.distributed initialised (it carries the NCCL id)")
        t = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            check(lib.ab_dist_unique_id(idbuf))
            t = torch.frombuffer(bytearray(bytes(idbuf)), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            t = t.cuda(device)
        dist.broadcast(t, src=0)
        raw = bytes(t.cpu().numpy().tobytes())
        idbuf = (C.c_ubyte * 128).from_buffer_copy(raw)
    check(lib.ab_dist_init(rank, world, idbuf))
    _initialized, _rank, _world = True, rank, world


def mpi_finalize() -> None:
    global _initialized
    check(lib.ab_dist_finalize())
    _initialized = False


def mpi_rank() -> int:
    return _rank


def mpi_size() -> int:
    return _world


def row_partition(N: int, world: int, rank: int, align: int = 1):
    """Contiguous [ilower, iupper] (inclusive) like Hypre's IJ layout (mpi.jl:392-401); `align`
    keeps block edges on multiples of a plane size for stencil grids."""
    units = N // align
    lo = (units * rank // world) * align
    hi = (units * (rank + 1) // world) * align if rank < world - 1 else N
    return lo, hi - 1


class mpi_SparseTensor:
    """Local row block of a distributed matrix (mpi.jl:420-474): rows ilower..iupper of an N x N
    matrix as a CSR handle with GLOBAL column ids."""

    def __init__(self, local: SparseTensor, ilower: int, iupper: int, N: int):
        self.local, self.ilower, self.iupper, self.N = local, int(ilower), int(iupper), int(N)
        dh = C.c_int64(0)
        check(lib.ab_dist_csr_create(local.handle(), self.ilower, self.iupper, self.N, C.byref(dh)))
        self._dh = dh.value

    @classmethod
    def poisson3d(cls, nx: int, ny: int, nz: int, rank: int | None = None, world: int | None = None):
        """z-slab partition of the 7-point Laplacian, built direct-to-CSR on this rank's GPU."""
        rank = _rank if rank is None else rank
        world = _world if world is None else world
        N = nx * ny * nz
        lo, hi = row_partition(N, world, rank, align=nx * ny)
        h = C.c_int64(0)
        check(lib.ab_poisson3d_csr(nx, ny, nz, lo, hi + 1, C.byref(h)))
        return cls(SparseTensor.from_handle(h.value), lo, hi, N)

    @classmethod
    def poisson3d_kappa(cls, kappa, nx: int, ny: int, nz: int, rank: int | None = None, world: int | None = None):
        """Same partition, variable coefficients: kappa is the GLOBAL cell field [nx*ny*nz] (host or device)."""
        rank = _rank if rank is None else rank
        world = _world if world is None else world
        N = nx * ny * nz
        lo, hi = row_partition(N, world, rank, align=nx * ny)
        kk, pk = as_array(kappa, np.float64)
        h = C.c_int64(0)
        check(lib.ab_poisson3d_kappa_csr(nx, ny, nz, pk, lo, hi + 1, C.byref(h)))
        return cls(SparseTensor.from_handle(h.value), lo, hi, N)

    @property
    def nloc(self) -> int:
        return self.iupper - self.ilower + 1

    def __matmul__(self, x_local):
        kx, px = as_array(x_local, np.float64)
        y = empty_like_kind(kx, (self.nloc,), np.float64)
        check(lib.ab_dist_spmv(self._dh, px, ptr(y)))
        return y

    def solve(self, b_local, solver: str | None = None, tol: float = 0.0, maxit: int = 0):
        """mpi_SparseTensor \\ b (mpi.jl:493-500); returns (u_local, iters, relres)."""
        solver = options.mpi.solver if solver is None else solver
        kb, pb = as_array(b_local, np.float64)
        u = empty_like_kind(kb, (self.nloc,), np.float64)
        it, rr = C.c_int(0), C.c_double(0)
        check(lib.ab_dist_solve(self._dh, solver.encode(), pb, ptr(u), tol, maxit, C.byref(it), C.byref(rr)))
        return u, it.value, rr.value

    def solve_grad(self, grad_u_local, u_local, tol: float = 0.0, maxit: int = 0):
        """Backward of `mpi_SparseTensor \\ b`: (grad of the local CSR values, per stored entry of this rank's
        rows; grad_b_local).  Collective — every rank calls it."""
        kg, pg = as_array(grad_u_local, np.float64)
        ku, pu = as_array(u_local, np.float64)
        nnz = self.nnz_local
        gv = empty_like_kind(kg, (nnz,), np.float64)
        gf = empty_like_kind(kg, (self.nloc,), np.float64)
        check(lib.ab_dist_solve_grad(self._dh, pg, pu, ptr(gv), ptr(gf), tol, maxit, None, None))
        return gv, gf

    def adjoint(self):
        """adjoint(A::mpi_SparseTensor) (mpi.jl:544-563): A' under the same row partition.  Collective."""
        dh = C.c_int64(0)
        check(lib.ab_dist_transpose(self._dh, C.byref(dh)))
        return mpi_SparseTensor._from_dist_handle(dh.value, self.ilower, self.iupper, self.N)

    @property
    def T(self):
        return self.adjoint()

    def coo(self):
        """This rank's rows as (idx2 [nnz,2] int64: local row, GLOBAL column; values) — mpi_get_matrix (MPITensor.h:39-52)."""
        nnz = self.nnz_local
        idx2, vals = np.empty((nnz, 2), np.int64), np.empty(nnz, np.float64)
        check(lib.ab_dist_export_coo(self._dh, ptr(idx2), ptr(vals)))
        return idx2, vals

    def to_SparseTensor(self) -> SparseTensor:
        """SparseTensor(sp::mpi_SparseTensor) (mpi.jl:502-504): the local row block, nloc x N."""
        idx2, vals = self.coo()
        return SparseTensor(idx2[:, 0] + 1, idx2[:, 1] + 1, vals, self.nloc, self.N)

    @classmethod
    def _from_dist_handle(cls, dh: int, ilower: int, iupper: int, N: int):
        self = cls.__new__(cls)
        self.local, self.ilower, self.iupper, self.N, self._dh = None, int(ilower), int(iupper), int(N), int(dh)
        return self

    @property
    def nnz_local(self) -> int:
        if self.local is not None:
            return self.local.query()[2]
        n = C.c_int64(0)
        check(lib.ab_dist_query(self._dh, None, None, C.byref(n)))
        return n.value

    def cg_fixed(self, b_local, iters: int):
        kb, pb = as_array(b_local, np.float64)
        u = empty_like_kind(kb, (self.nloc,), np.float64)
        rr = C.c_double(0)
        check(lib.ab_dist_cg_fixed(self._dh, pb, ptr(u), int(iters), C.byref(rr)))
        return u, rr.value

    def close(self) -> None:
        if getattr(self, "_dh", None):
            lib.ab_dist_destroy(self._dh)
            self._dh = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
