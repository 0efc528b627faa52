"""load_op / load_op_and_grad / load_system_op over libadcme_b200
(/root/reference/src/extra.jl:134-260): an op `FooBar` is looked up by its snake_case name and
comes back as a Python callable; when the library has a `…Grad` companion the callable carries it
(`op.grad`), with the reference's gradient-op argument order
(grad of each float output…, outputs…, inputs…) -> one gradient per input.

Only the ops of the sparse hot path exist; asking for anything else raises, the way
a missing symbol does in the reference (extra.jl:197-235).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import sparse as _sp
from . import structure as _st
from ._lib import LIB_PATH, as_array, check, empty_like_kind, lib, ptr

libadcme = LIB_PATH


def _n(x):
    return _sp._numel(x)


# ------------------------------------------------------------------ sparse_solver
def sparse_solver(ii, jj, vv, f, method="SparseLU"):
    """SparseSolver(ii,jj,vv,f,method) -> u (SparseSolver.cpp:21-45); ii/jj 1-based int64."""
    ki, pi = as_array(ii, np.int64); kj, pj = as_array(jj, np.int64)
    kv, pv = as_array(vv, np.float64); kf, pf = as_array(f, np.float64)
    u = empty_like_kind(kf, (_n(kf),), np.float64)
    check(lib.ab_sparse_solver(pi, pj, pv, _n(kv), pf, _n(kf), method.encode(), ptr(u)))
    return u


def sparse_solver_grad(grad_u, u, ii, jj, vv, f, method="SparseLU"):
    """SparseSolverGrad(grad_u,u,ii,jj,vv,f,method) -> (grad_ii,grad_jj,grad_vv,grad_f,grad_method);
    only grad_vv, grad_f are defined (SparseSolver.cpp:46-60, SparseSolver.h:46-70)."""
    kg, pg = as_array(grad_u, np.float64); ku, pu = as_array(u, np.float64)
    ki, pi = as_array(ii, np.int64); kj, pj = as_array(jj, np.int64)
    kv, pv = as_array(vv, np.float64); kf, pf = as_array(f, np.float64)
    gvv = empty_like_kind(kv, (_n(kv),), np.float64)
    gf = empty_like_kind(kf, (_n(kf),), np.float64)
    check(lib.ab_sparse_solver_grad(pg, pu, pi, pj, pv, _n(kv), pf, _n(kf), method.encode(), ptr(gvv), ptr(gf)))
    return None, None, gvv, gf, None


# ------------------------------------------------------------------ sparse_compress
def sparse_compress(indices, v):
    return _sp.compress_indices(indices, v)


def sparse_compress_grad(grad_nv, nindices, nv, indices, v):
    """SparseCompressGrad: gradients for (indices, v); indices get none (SparseCompress.cpp:40-166)."""
    return None, _sp.compress_grad(indices, grad_nv)


# ------------------------------------------------------------------ sparse accumulator trio
def sparse_accumulator(tol, n, handle):
    return _sp.SparseAssembler(int(handle), int(n), float(tol))


def sparse_accumulator_add(handle, row, cols, vals):
    return _sp.accumulate(int(handle), int(row), cols, vals)


def sparse_accumulator_copy(ops):
    """-> (rows, cols, vals) int32/int32/f64, 1-based, handle destroyed (SparseAccumulator.cpp:163-230)."""
    hs = np.atleast_1d(np.asarray(ops)).astype(np.int64)
    h = int(hs[0])
    k = lib.ab_asm_nnz(h)
    if k < 0:
        check(int(k))
    r, c, v = np.empty(k, np.int32), np.empty(k, np.int32), np.empty(k, np.float64)
    check(lib.ab_asm_copy(h, ptr(r), ptr(c), ptr(v)))
    return r, c, v


# ------------------------------------------------------------------ factorization / solve
def sparse_factorization(ii, jj, vv, d, max_cache_size):
    ki, pi = as_array(ii, np.int64); kj, pj = as_array(jj, np.int64); kv, pv = as_array(vv, np.float64)
    o = C.c_int64(0)
    check(lib.ab_factorize(pi, pj, pv, _n(kv), int(d), int(max_cache_size), C.byref(o)))
    return o.value


def solve(rhs, ii, jj, vv, o):
    kr, pr = as_array(rhs, np.float64)
    out = empty_like_kind(kr, (_n(kr),), np.float64)
    check(lib.ab_factorized_solve(int(o), pr, _n(kr), ptr(out)))
    return out


def solve_grad(grad_out, out, rhs, ii, jj, vv, o):
    """SolveGrad -> (grad_rhs, grad_ii, grad_jj, grad_vv, grad_o) (Solve.cpp:45-60, Solve.h:15-33)."""
    kg, pg = as_array(grad_out, np.float64); ko, po = as_array(out, np.float64)
    ki, pi = as_array(ii, np.int64); kj, pj = as_array(jj, np.int64)
    d, nv = _n(ko), _n(ki)
    grhs = empty_like_kind(kg, (d,), np.float64)
    gvv = empty_like_kind(kg, (nv,), np.float64)
    check(lib.ab_factorized_solve_grad(int(o), pg, po, pi, pj, nv, d, ptr(grhs), ptr(gvv)))
    return grhs, None, None, gvv, None


# ------------------------------------------------------------------ mpi_create_matrix
def mpi_create_matrix(indices, values, ilower, iupper):
    """MPICreateMatrix (MPITensor.cpp:23-138, MPITensor.h:1-36) -> rows, ncols, cols, values."""
    ki, pi = as_array(indices, np.int64)
    n = _n(ki) // 2
    nrow = C.c_int64(0)
    check(lib.ab_coo_to_rows_count(n, pi, C.byref(nrow)))
    rows = empty_like_kind(ki, (nrow.value,), np.int32)
    ncols = empty_like_kind(ki, (nrow.value,), np.int32)
    cols = empty_like_kind(ki, (n,), np.int32)
    check(lib.ab_coo_to_rows(n, pi, int(ilower), ptr(rows), ptr(ncols), ptr(cols)))
    return rows, ncols, cols, values


_OPS = {
    "sparse_solver": (sparse_solver, sparse_solver_grad),
    "sparse_compress": (sparse_compress, sparse_compress_grad),
    "sparse_accumulator": (sparse_accumulator, None),
    "sparse_accumulator_add": (sparse_accumulator_add, None),
    "sparse_accumulator_copy": (sparse_accumulator_copy, None),
    "sparse_factorization": (sparse_factorization, None),
    "solve": (solve, solve_grad),
    "mpi_create_matrix": (mpi_create_matrix, None),
    "zero_out_row": (_st.zero_out_row_op, _st.zero_out_row_grad),
    "sparse_scatter_update": (_st.sparse_scatter_update_op, _st.sparse_scatter_update_grad),
    "sparse_indexing": (_st.sparse_indexing_op, _st.sparse_indexing_grad),
    "sparse_least_square": (_st.sparse_least_square_op, _st.sparse_least_square_grad),
    "tri_solve": (_st.tri_solve_op, _st.tri_solve_grad),
    "sparse_concate": (_st.sparse_concate_op, _st.sparse_concate_grad),
    "sparse_sparse_mat_mul": (_st.sparse_sparse_mat_mul_op, None),
    "diag_sparse_mat_mul": (_st.diag_sparse_mat_mul_op, None),
    "sparse_diag_mat_mul": (_st.sparse_diag_mat_mul_op, None),
}


class _Op:
    def __init__(self, name, fwd, grad):
        self.__name__, self._fwd, self.grad = name, fwd, grad

    def __call__(self, *a, **k):
        return self._fwd(*a, **k)


def load_op(oplibpath, opname: str):
    """extra.jl:134-158."""
    if opname not in _OPS:
        raise KeyError(f"op '{opname}' is not exported by {oplibpath} (hot-path ops: {sorted(_OPS)})")
    return _Op(opname, _OPS[opname][0], None)


def load_op_and_grad(oplibpath, opname: str, multiple: bool = False):
    """extra.jl:160-235.  `multiple` is accepted for signature parity (ops here return tuples natively)."""
    if opname not in _OPS:
        raise KeyError(f"op '{opname}' is not exported by {oplibpath} (hot-path ops: {sorted(_OPS)})")
    fwd, grad = _OPS[opname]
    return _Op(opname, fwd, grad)


def load_system_op(opname: str, grad: bool = True, multiple: bool = False):
    """extra.jl:250-260."""
    return load_op_and_grad(libadcme, opname, multiple=multiple) if grad else load_op(libadcme, opname)


def get_library_symbols(path: str = LIB_PATH):
    """extra.jl:1023-1034: exported symbol names of a shared library."""
    import subprocess

    out = subprocess.run(["nm", "-D", "--defined-only", path], check=True, capture_output=True, text=True).stdout
    return sorted(line.split()[-1] for line in out.splitlines() if line.strip())
