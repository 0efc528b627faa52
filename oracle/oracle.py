"""Python face of the CPU oracle (TEST INFRASTRUCTURE — see the header of adcme_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  It wraps oracle/liboracle.so (the C restatement of the reference's algorithms) with
ctypes and adds a SuperLU-based direct solve for mid-size systems: Eigen::SparseLU, which the
reference calls at deps/CustomOps/SparseSolver/SparseSolver.h:13-44, is the same supernodal LU
family as SuperLU, and parity is on the solution (<= 1e-10 relative residual), not on pivots.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "adcme_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_accum_create.restype = C.c_void_p
        _lib.orc_accum_create.argtypes = [C.c_int, C.c_double]
        _lib.orc_accum_add.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        _lib.orc_accum_nnz.restype = C.c_int64
        _lib.orc_accum_nnz.argtypes = [C.c_void_p]
        _lib.orc_accum_copy.argtypes = [C.c_void_p] * 4
        _lib.orc_compress.restype = C.c_int64
        _lib.orc_coo_to_rows_count.restype = C.c_int64
        _lib.orc_coo_to_csr.restype = C.c_int64
        _lib.orc_poisson2d_matrix.restype = C.c_int64
        _lib.orc_pcg_jacobi.restype = C.c_int
        _lib.orc_pcg_jacobi.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_double, C.c_int, C.c_void_p, C.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ---------------------------------------------------------------- SparseAssembler
class Accumulator:
    """SparseAccum (deps/CustomOps/SparseAccumulate/Impl.cpp:11-48)."""

    def __init__(self, nrow: int, tol: float = 0.0):
        self._h = lib().orc_accum_create(int(nrow), float(tol))

    def add(self, row: int, cols, vals) -> int:
        cols, vals = _i32(cols), _f64(vals)
        return lib().orc_accum_add(self._h, int(row), _p(cols), _p(vals), len(vals))

    def nnz(self) -> int:
        return int(lib().orc_accum_nnz(self._h))

    def copy(self):
        k = self.nnz()
        rows, cols, vals = np.zeros(k, np.int32), np.zeros(k, np.int32), np.zeros(k, np.float64)
        lib().orc_accum_copy(self._h, _p(rows), _p(cols), _p(vals))
        self._h = None
        return rows, cols, vals


# ---------------------------------------------------------------- compress / coo_to_rows
def compress(idx2, v):
    idx2, v = _i64(idx2).reshape(-1, 2), _f64(v)
    N = len(v)
    oi, ov = np.zeros((max(N, 1), 2), np.int64), np.zeros(max(N, 1), np.float64)
    k = lib().orc_compress(C.c_int64(N), _p(idx2), _p(v), _p(oi), _p(ov))
    return oi[:k].copy(), ov[:k].copy()


def compress_grad(idx2, grad_nv):
    idx2, g = _i64(idx2).reshape(-1, 2), _f64(grad_nv)
    N = idx2.shape[0]
    out = np.zeros(N, np.float64)
    lib().orc_compress_grad(C.c_int64(N), _p(idx2), _p(g), _p(out))
    return out


def coo_to_rows(idx2, ilower: int):
    idx2 = _i64(idx2).reshape(-1, 2)
    n = idx2.shape[0]
    nrow = lib().orc_coo_to_rows_count(C.c_int64(n), _p(idx2))
    rows, ncols, cols = np.zeros(nrow, np.int32), np.zeros(nrow, np.int32), np.zeros(n, np.int32)
    lib().orc_coo_to_rows(C.c_int64(n), _p(idx2), C.c_int64(ilower), _p(rows), _p(ncols), _p(cols))
    return rows, ncols, cols


# ---------------------------------------------------------------- COO -> CSR (setFromTriplets)
def coo_to_csr(ii, jj, vv, m: int, n: int, base: int = 1):
    ii, jj, vv = _i64(ii), _i64(jj), _f64(vv)
    T = len(vv)
    rowptr = np.zeros(m + 1, np.int32)
    colind = np.zeros(max(T, 1), np.int32)
    values = np.zeros(max(T, 1), np.float64)
    slot = np.zeros(max(T, 1), np.int32)
    nnz = lib().orc_coo_to_csr(C.c_int64(T), _p(ii), _p(jj), _p(vv), C.c_int(base), C.c_int64(m), C.c_int64(n),
                               _p(rowptr), _p(colind), _p(values), _p(slot))
    if nnz < 0:
        raise IndexError("triplet index out of range")
    return rowptr, colind[:nnz].copy(), values[:nnz].copy(), slot[:T].copy()


# ---------------------------------------------------------------- SpMM + grads (TF semantics)
def spmm_coo(idx2, val, m: int, B, adjoint_a: bool = False):
    idx2, val = _i64(idx2).reshape(-1, 2), _f64(val)
    B = _f64(B)
    vec = B.ndim == 1
    B2 = B.reshape(len(B), -1)
    k = B2.shape[1]
    out = np.zeros((m, k), np.float64)
    lib().orc_spmm_coo(C.c_int64(len(val)), _p(idx2), _p(val), C.c_int64(m), C.c_int64(k), _p(B2), _p(out),
                       C.c_int(1 if adjoint_a else 0))
    return out[:, 0].copy() if vec else out


def spmm_grad_values(idx2, grad, B):
    idx2 = _i64(idx2).reshape(-1, 2)
    grad, B = _f64(grad), _f64(B)
    g2, B2 = grad.reshape(len(grad), -1), B.reshape(len(B), -1)
    out = np.zeros(idx2.shape[0], np.float64)
    lib().orc_spmm_grad_values(C.c_int64(idx2.shape[0]), _p(idx2), C.c_int64(g2.shape[1]), _p(g2), _p(B2), _p(out))
    return out


def csr_spmm(rowptr, colind, values, X):
    rowptr, colind, values = _i32(rowptr), _i32(colind), _f64(values)
    X = _f64(X)
    vec = X.ndim == 1
    X2 = X.reshape(len(X), -1)
    m, k = len(rowptr) - 1, X2.shape[1]
    Y = np.zeros((m, k), np.float64)
    lib().orc_csr_spmm(C.c_int64(m), _p(rowptr), _p(colind), _p(values), C.c_int64(k), _p(X2), _p(Y))
    return Y[:, 0].copy() if vec else Y


# ---------------------------------------------------------------- SparseSolver fwd / bwd
class SolverError(RuntimeError):
    """Mirrors Status(INTERNAL, "Sparse solver factorization failed.") (SparseSolver.cpp:119-138)."""


def sparse_solver(ii, jj, vv, f, d: int | None = None, dense_limit: int = 600):
    ii, jj, vv, f = _i64(ii), _i64(jj), _f64(vv), _f64(f)
    d = len(f) if d is None else d
    if d <= dense_limit:
        u = np.zeros(d, np.float64)
        rc = lib().orc_sparse_solver(_p(ii), _p(jj), _p(vv), C.c_int64(len(vv)), _p(f), C.c_int64(d), C.c_int(0), _p(u))
        if rc != 0:
            raise SolverError("Sparse solver factorization failed.")
        return u
    return _splu_solve(ii, jj, vv, f, d, transpose=False)


def sparse_solver_grad(grad_u, u, ii, jj, vv, d: int | None = None, dense_limit: int = 600):
    ii, jj, vv, grad_u, u = _i64(ii), _i64(jj), _f64(vv), _f64(grad_u), _f64(u)
    d = len(u) if d is None else d
    if d <= dense_limit:
        gv, gf = np.zeros(len(vv), np.float64), np.zeros(d, np.float64)
        rc = lib().orc_sparse_solver_grad(_p(grad_u), _p(u), _p(ii), _p(jj), _p(vv), C.c_int64(len(vv)), C.c_int64(d),
                                          _p(gv), _p(gf))
        if rc != 0:
            raise SolverError("Sparse solver factorization failed.")
        return gv, gf
    x = _splu_solve(ii, jj, vv, grad_u, d, transpose=True)
    gv = np.zeros(len(vv), np.float64)
    lib().orc_solver_grad_vv(_p(x), _p(u), _p(ii), _p(jj), C.c_int64(len(vv)), _p(gv))
    return gv, x


def _splu_solve(ii, jj, vv, f, d, transpose):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    r, c = (jj - 1, ii - 1) if transpose else (ii - 1, jj - 1)
    A = sp.coo_matrix((vv, (r, c)), shape=(d, d)).tocsc()   # duplicates summed, like setFromTriplets
    try:
        lu = spla.splu(A)
    except RuntimeError as e:   # exactly singular
        raise SolverError("Sparse solver factorization failed.") from e
    x = lu.solve(f)
    # one step of iterative refinement keeps the oracle itself well below the 1e-10 gate
    x = x + lu.solve(f - A @ x)
    return x


# ---------------------------------------------------------------- structural ops
# Plain-Python restatements (small cases only) of the reference loops, line by line.
def zero_out_row(indices, vv, bd):
    """ZeroOutRow::forward + move_data (deps/CustomOps/ZeroOutRow/ZeroOutRow.h:19-48): indices are the
    0-based [N,2] matrix, bd is 1-based; entries whose row is in {bd-1} are skipped, order kept."""
    indices, vv = _i64(indices).reshape(-1, 2), _f64(vv)
    bdset = {int(b) - 1 for b in np.asarray(bd).ravel()}
    keep = [i for i in range(len(vv)) if int(indices[i, 0]) not in bdset]
    return indices[keep].copy().reshape(-1, 2), vv[keep].copy()


def zero_out_row_grad(indices, bd, grad_oval):
    """ZeroOutRow::backward (ZeroOutRow.h:50-57); dropped slots (left unwritten there) are 0 here."""
    indices, g = _i64(indices).reshape(-1, 2), _f64(grad_oval)
    bdset = {int(b) - 1 for b in np.asarray(bd).ravel()}
    out, k = np.zeros(indices.shape[0]), 0
    for i in range(indices.shape[0]):
        if int(indices[i, 0]) in bdset:
            continue
        out[i] = g[k]
        k += 1
    return out


def sparse_scatter_update(oii, ojj, ovv, uii, ujj, uvv, m, n, ii, jj):
    """forward (deps/CustomOps/SparseScatterUpdate/SparseScatterUpdate.h:20-37), all 1-based."""
    oii, ojj, ovv = _i64(oii), _i64(ojj), _f64(ovv)
    uii, ujj, uvv = _i64(uii), _i64(ujj), _f64(uvv)
    ii, jj = _i64(ii), _i64(jj)
    iset, jset = set(ii.tolist()), set(jj.tolist())
    I, J, V = [], [], []
    for i in range(len(ovv)):
        if int(oii[i]) in iset and int(ojj[i]) in jset:
            continue
        I.append(oii[i]); J.append(ojj[i]); V.append(ovv[i])
    for i in range(len(uvv)):
        I.append(ii[uii[i] - 1]); J.append(jj[ujj[i] - 1]); V.append(uvv[i])
    return np.array(I, np.int64), np.array(J, np.int64), np.array(V, np.float64)


def sparse_scatter_update_grad(grad_out, oii, ojj, un, ii, jj):
    """backward (SparseScatterUpdate.h:39-60) -> (grad_ovv, grad_uvv)."""
    g, oii, ojj = _f64(grad_out), _i64(oii), _i64(ojj)
    iset, jset = set(_i64(ii).tolist()), set(_i64(jj).tolist())
    govv, k0 = np.zeros(len(oii)), 0
    for i in range(len(oii)):
        if int(oii[i]) in iset and int(ojj[i]) in jset:
            govv[i] = 0.0
        else:
            govv[i] = g[k0]; k0 += 1
    return govv, g[k0:k0 + un].copy()


def sparse_concate(ii1, jj1, vv1, m1, n1, ii2, jj2, vv2, hcat):
    """Forward (deps/CustomOps/SparseConcate/SparseConcate.h:11-27)."""
    ii1, jj1, ii2, jj2 = _i64(ii1), _i64(jj1), _i64(ii2), _i64(jj2)
    if hcat:
        return np.concatenate([ii1, ii2]), np.concatenate([jj1, jj2 + n1]), np.concatenate([_f64(vv1), _f64(vv2)])
    return np.concatenate([ii1, ii2 + m1]), np.concatenate([jj1, jj2]), np.concatenate([_f64(vv1), _f64(vv2)])


def sparse_to_dense(ij, vv, m, n):
    """forward (deps/CustomOps/SparseToDense/SparseToDense.h:1-6): A[i*n+j] += v in input order."""
    ij, vv = _i64(ij).reshape(-1, 2), _f64(vv)
    A = np.zeros(m * n)
    for k in range(len(vv)):
        A[ij[k, 0] * n + ij[k, 1]] += vv[k]
    return A.reshape(m, n)


def sparse_least_square(ii, jj, vv, ff, m, n):
    """forward (deps/CustomOps/SparseLeastSquare/SparseLeastSquare.h:10-32): u = argmin ||A u - f|| per batch
    row (the reference factors A with SparseQR; a dense QR-based lstsq gives the same minimiser)."""
    A = np.zeros((m, n))
    np.add.at(A, (_i64(ii) - 1, _i64(jj) - 1), _f64(vv))              # setFromTriplets sums duplicates
    F = _f64(ff).reshape(-1, m)
    U = np.stack([np.linalg.lstsq(A, F[l], rcond=None)[0] for l in range(F.shape[0])])
    return U[0] if np.ndim(ff) == 1 else U


def sparse_least_square_grad(grad_u, u, ii, jj, vv, ff, m, n):
    """backward (SparseLeastSquare.h:34-83), the loops as written: M = A'A, x = M^-1 g, grad_f = A x,
    grad_vv[k] += -t[i_k] x[j_k] + f[i_k] x[j_k] - u[j_k] * sum_r At(r, i_k) x[r]   (t = A u)."""
    ii, jj = _i64(ii) - 1, _i64(jj) - 1
    A = np.zeros((m, n))
    np.add.at(A, (ii, jj), _f64(vv))
    M = A.T @ A
    G, Uu, F = _f64(grad_u).reshape(-1, n), _f64(u).reshape(-1, n), _f64(ff).reshape(-1, m)
    gvv, gfs = np.zeros(len(ii)), []
    for l in range(G.shape[0]):
        t = A @ Uu[l]
        x = np.linalg.solve(M, G[l])
        gf = A @ x
        gfs.append(gf)
        for k in range(len(ii)):
            gvv[k] += -t[ii[k]] * x[jj[k]]
            gvv[k] += F[l][ii[k]] * x[jj[k]]
            gvv[k] += -Uu[l][jj[k]] * float(A[ii[k], :] @ x)          # InnerIterator over column i_k of At
    gfs = np.stack(gfs)
    return gvv, (gfs[0] if np.ndim(ff) == 1 else gfs)


def tri_solve(a, b, c, d):
    """TriSolve_forward (deps/CustomOps/TriSolve/TriSolve.h:1-20): the Thomas algorithm, line by line."""
    a, b, c, d = _f64(a), _f64(b), _f64(c), _f64(d)
    n = len(b)
    B, D = b.copy(), d.copy()
    for i in range(1, n):
        w = a[i - 1] / B[i - 1]
        B[i] = B[i] - w * c[i - 1]
        D[i] = D[i] - w * D[i - 1]
    X = np.zeros(n)
    if n:
        X[n - 1] = D[n - 1] / B[n - 1]
    for i in range(n - 2, -1, -1):
        X[i] = (D[i] - c[i] * X[i + 1]) / B[i]
    return X


def tri_solve_grad(grad_x, x, a, b, c):
    """TriSolve_backward (TriSolve.h:22-37) -> (grad_a, grad_b, grad_c, grad_d)."""
    x = _f64(x)
    n = len(x)
    gd = tri_solve(c, b, a, grad_x)
    ga, gb, gc = np.zeros(max(n - 1, 0)), np.zeros(n), np.zeros(max(n - 1, 0))
    for i in range(n):
        if i > 0:
            ga[i - 1] = -gd[i] * x[i - 1]
        gb[i] = -gd[i] * x[i]
        if i < n - 1:
            gc[i] = -gd[i] * x[i + 1]
    return ga, gb, gc, gd


def _csc_from_triplets(ii, jj, vv, m, n, base):
    """Eigen setFromTriplets into a column-major matrix = our CSR restatement with the roles of the
    two indices swapped: -> (colptr, rowind, values), rows sorted, duplicates summed in input order."""
    cp, ri, va, _ = coo_to_csr(jj, ii, vv, n, m, base=base)
    return cp, ri, va


def sparse_indexing(ii1, jj1, vv1, m, n, ii, jj):
    """forward (deps/CustomOps/SparseIndexing/SparseIndexing.h:23-63): 1-based in, 1-based block
    positions out, ordered by position in jj then ascending row."""
    ii, jj = _i64(ii), _i64(jj)
    ii_map = {int(v): p for p, v in enumerate(ii.tolist())}      # later duplicates overwrite
    jj_map = {int(v): p for p, v in enumerate(jj.tolist())}
    cp, ri, va = _csc_from_triplets(ii1, jj1, vv1, m, n, base=1)
    ii_sorted = sorted(ii.tolist())
    I, J, V = [], [], []
    for k in range(len(jj)):
        col = int(jj[k])
        rows = {int(ri[q]) + 1: va[q] for q in range(cp[col - 1], cp[col])}
        for r in ii_sorted:                       # std::set_intersection of two sorted lists
            if r in rows:
                I.append(ii_map[r] + 1); J.append(jj_map[col] + 1); V.append(rows[r])
    return np.array(I, np.int64), np.array(J, np.int64), np.array(V, np.float64)


def sparse_indexing_grad(grad_vv0, ii0, jj0, ii1, jj1, ii, jj):
    """backward (SparseIndexing.h:66-83): the std::map keeps the LAST triplet index of each (i,j)."""
    g, ii0, jj0 = _f64(grad_vv0), _i64(ii0), _i64(jj0)
    ii1, jj1, ii, jj = _i64(ii1), _i64(jj1), _i64(ii), _i64(jj)
    mp = {(int(a), int(b)): k for k, (a, b) in enumerate(zip(ii1.tolist(), jj1.tolist()))}
    out = np.zeros(len(ii1))
    for i in range(len(g)):
        out[mp[(int(ii[ii0[i] - 1]), int(jj[jj0[i] - 1]))]] += g[i]
    return out


def sparse_sparse_mat_mul(ii1, jj1, vv1, ii2, jj2, vv2, m, n, k):
    """forward (deps/CustomOps/SparseMatMul/SparseMatMul.h:11-32) + the emission loop of
    SparseMatMul.cpp: 0-based triplets in, C = A*B by Eigen's column-major conservative product
    (for each column j of B, for each stored B(kk,j) in ascending kk, for each stored A(i,kk):
    C(i,j) += A(i,kk)*B(kk,j)), 1-based column-major triplets out, structural zeros kept."""
    acp, ari, ava = _csc_from_triplets(ii1, jj1, vv1, m, n, base=0)
    bcp, bri, bva = _csc_from_triplets(ii2, jj2, vv2, n, k, base=0)
    I, J, V = [], [], []
    for j in range(k):
        acc = {}
        for q in range(bcp[j], bcp[j + 1]):
            kk, y = int(bri[q]), bva[q]
            for p in range(acp[kk], acp[kk + 1]):
                i, x = int(ari[p]), ava[p]
                acc[i] = acc[i] + x * y if i in acc else x * y
        for i in sorted(acc):
            I.append(i + 1); J.append(j + 1); V.append(acc[i])
    return np.array(I, np.int64), np.array(J, np.int64), np.array(V, np.float64)


def _colmajor_from_triplets(ii, jj, vv, m, k):
    cp, ri, va = _csc_from_triplets(ii, jj, vv, m, k, base=0)
    J = np.repeat(np.arange(k, dtype=np.int64), np.diff(cp)) + 1
    return ri.astype(np.int64) + 1, J, va


def diag_sparse_mat_mul(vv1, ii2, jj2, vv2, m, k):
    """forward_diag_sparse (SparseMatMul.h:34-42): triplets (ii2, jj2, vv2*vv1[ii2]) -> setFromTriplets."""
    ii2, jj2 = _i64(ii2), _i64(jj2)
    return _colmajor_from_triplets(ii2, jj2, _f64(vv2) * _f64(vv1)[ii2], m, k)


def sparse_diag_mat_mul(ii1, jj1, vv1, vv2, m, k):
    """forward_sparse_diag (SparseMatMul.h:44-52): triplets (ii1, jj1, vv1*vv2[jj1])."""
    ii1, jj1 = _i64(ii1), _i64(jj1)
    return _colmajor_from_triplets(ii1, jj1, _f64(vv1) * _f64(vv2)[jj1], m, k)


# ---------------------------------------------------------------- Newton-Raphson
def newton_raphson(func, u0, theta=None, max_iter=100, tol=1e-12, rtol=1e-12, LM=0.0):
    """Restatement of the newton_raphson while-loop (src/optim.jl:462-592) with a direct solve
    (scipy SuperLU standing in for Eigen SparseLU).  func(theta, u) -> (r, J) with J a scipy sparse
    matrix.  Returns (x, res, U, converged, iter) exactly as NRResult is filled (optim.jl:560-591):
    ta_r[i] holds ||r(u_{i-2})||, the loop continues while ta_r[i-1] >= tol and ta_r[i-2] >= rtol and
    i <= max_iter, the answer is ta_u[i_-1]."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    u = _f64(u0).copy()
    r0, _ = func(theta, u)
    tol0 = float(np.linalg.norm(r0))
    ta_r, ta_u, i = {1: tol0}, {1: u}, 2
    while True:
        cont = (i < max_iter + 1) if i == 2 else (ta_r[i - 1] >= tol and ta_r[i - 2] >= rtol and i <= max_iter)
        if not cont:
            break
        u_ = ta_u[i - 1]
        r_, J = func(theta, u_)
        J = sp.csc_matrix(J)
        if LM > 0.0:
            J = J + LM * sp.identity(J.shape[0], format="csc")       # optim.jl:520-524
        delta = spla.splu(J).solve(_f64(r_))
        ta_r[i] = float(np.linalg.norm(r_))
        ta_u[i] = u_ - delta
        i += 1
    if tol0 < tol:
        return u, np.array([tol0]), u.reshape(-1, 1), i < max_iter, 1
    res = np.array([ta_r[k] for k in range(2, i)])
    U = np.stack([ta_u[k] for k in range(1, i - 1)], axis=1)
    return ta_u[i - 1], res, U, i < max_iter, i - 2


def newton_adjoint_grad(func_r_theta_jac, A, dy):
    """newton_raphson_with_grad backward (optim.jl:628-638): g = A'^-1 dy ; grad_theta = -(dr/dtheta)' g.
    `func_r_theta_jac` is the dense/sparse Jacobian of r with respect to theta at the solution."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    g = spla.splu(sp.csc_matrix(A).T.tocsc()).solve(_f64(dy))
    return -(func_r_theta_jac.T @ g)


# ---------------------------------------------------------------- Poisson generators
def poisson2d_matrix(n: int, kext):
    """GetPoissonMatrixForward (GetPoissonMatrix.h:15-51), single-rank colext."""
    kext = _f64(kext).reshape(n + 2, n + 2)
    cols, val, ncols = np.zeros(5 * n * n, np.int64), np.zeros(5 * n * n, np.float64), np.zeros(n * n, np.int32)
    p = lib().orc_poisson2d_matrix(C.c_int64(n), _p(kext), _p(cols), _p(val), _p(ncols))
    return cols[:p].copy(), val[:p].copy(), ncols


def poisson2d_grad(n: int, xgrad, uext):
    xgrad, uext = _f64(xgrad), _f64(uext)
    out = np.zeros((n + 2) * (n + 2), np.float64)
    lib().orc_poisson2d_grad(C.c_int64(n), _p(xgrad), _p(uext), _p(out))
    return out


def poisson3d_csr(nx: int, ny: int, nz: int):
    N = nx * ny * nz
    nnz = C.c_int64(0)
    lib().orc_poisson3d_csr_count(C.c_int64(nx), C.c_int64(ny), C.c_int64(nz), C.byref(nnz))
    rowptr, colind, values = np.zeros(N + 1, np.int32), np.zeros(nnz.value, np.int32), np.zeros(nnz.value, np.float64)
    lib().orc_poisson3d_csr(C.c_int64(nx), C.c_int64(ny), C.c_int64(nz), _p(rowptr), _p(colind), _p(values))
    return rowptr, colind, values


# ---------------------------------------------------------------- CPU Jacobi-PCG (baseline port)
def num_threads() -> int:
    return int(lib().orc_num_threads())


def set_num_threads(n: int) -> int:
    """OpenMP threads of the CPU baseline legs; torchrun exports OMP_NUM_THREADS=1, so bench.py sets this explicitly."""
    lib().orc_set_num_threads(int(n))
    return num_threads()


def pcg_jacobi(rowptr, colind, values, b, tol: float = 1e-12, maxit: int = 10000):
    rowptr, colind, values, b = _i32(rowptr), _i32(colind), _f64(values), _f64(b)
    n = len(b)
    x, work = np.zeros(n, np.float64), np.empty(4 * n, np.float64)
    relres = C.c_double(0)
    it = lib().orc_pcg_jacobi(C.c_int64(n), _p(rowptr), _p(colind), _p(values), _p(b), _p(x), float(tol), int(maxit),
                              C.byref(relres), _p(work))
    return x, it, relres.value


# ---------------------------------------------------------------- synthetic inputs shared by tests and bench
def facewise_triplets_2d(n: int, seed: int = 233, shuffle: bool = True):
    """FEM-style face-wise assembly stream of the 2-D 5-point Dirichlet Laplacian on n x n
    (the reference idiom is docs/src/tu_fem.md:41-54): every interior face emits
    (i,i,+1),(j,j,+1),(i,j,-1),(j,i,-1), every boundary face emits (i,i,+1).  1-based int64."""
    idx = np.arange(n * n, dtype=np.int64).reshape(n, n)
    I, J, V = [], [], []

    def faces(a, b):
        a, b = a.ravel(), b.ravel()
        I.extend([a, b, a, b]); J.extend([a, b, b, a])
        V.extend([np.ones(len(a)), np.ones(len(a)), -np.ones(len(a)), -np.ones(len(a))])

    faces(idx[:, :-1], idx[:, 1:])
    faces(idx[:-1, :], idx[1:, :])
    for edge in (idx[0, :], idx[-1, :], idx[:, 0], idx[:, -1]):
        I.append(edge.ravel()); J.append(edge.ravel()); V.append(np.ones(edge.size))
    ii, jj, vv = np.concatenate(I) + 1, np.concatenate(J) + 1, np.concatenate(V)
    if shuffle:
        p = np.random.default_rng(seed).permutation(len(vv))
        ii, jj, vv = ii[p], jj[p], vv[p]
    return ii, jj, vv


def facewise_triplets_3d(n: int, seed: int = 233, shuffle: bool = True):
    """Same scheme for the 3-D 7-point Laplacian on n^3: diag 6, off -1."""
    idx = np.arange(n ** 3, dtype=np.int64).reshape(n, n, n)
    I, J, V = [], [], []

    def faces(a, b):
        a, b = a.ravel(), b.ravel()
        I.extend([a, b, a, b]); J.extend([a, b, b, a])
        V.extend([np.ones(len(a)), np.ones(len(a)), -np.ones(len(a)), -np.ones(len(a))])

    faces(idx[:, :, :-1], idx[:, :, 1:])
    faces(idx[:, :-1, :], idx[:, 1:, :])
    faces(idx[:-1, :, :], idx[1:, :, :])
    for face in (idx[0], idx[-1], idx[:, 0], idx[:, -1], idx[:, :, 0], idx[:, :, -1]):
        I.append(face.ravel()); J.append(face.ravel()); V.append(np.ones(face.size))
    ii, jj, vv = np.concatenate(I) + 1, np.concatenate(J) + 1, np.concatenate(V)
    if shuffle:
        p = np.random.default_rng(seed).permutation(len(vv))
        ii, jj, vv = ii[p], jj[p], vv[p]
    return ii, jj, vv
