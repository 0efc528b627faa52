"""Host-side mirror of the structural SparseTensor ops of /root/reference/src/sparse.jl that sit
between assembly and solve and the remaining sparse algebra (N4), over the C-ABI
of csrc/structure.cu.  Nothing here computes; every function is one or two library calls.

    zero_out_row(A, bd)                  src/sparse.jl:790-797   -> ab_zero_out_row
    scatter_update / scatter_add         src/sparse.jl:340-386   -> ab_scatter_update
    getindex  A[i1, i2]                  src/sparse.jl:303-327   -> ab_sparse_indexing
    A + B, A - B                         src/sparse.jl:203-218   -> triplet concatenation / ab_csr_axpby
    A * B (SparseTensor x SparseTensor)  src/sparse.jl:579-595   -> ab_spgemm / ab_coo_scale
    spdiag, spzero                       src/sparse.jl:540-575, 631-658

Index conventions follow the reference: triplets and index lists are 1-based; `zero_out_row`'s
op-level `indices` are the 0-based [N,2] matrix of a tf.SparseTensor.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import as_array, check, empty_like_kind, lib, ptr


def _is_torch(x) -> bool:
    return _lib._is_torch(x)


def _n(x) -> int:
    return int(x.numel()) if _is_torch(x) else int(np.asarray(x).size)


def _fetch_trip(th: int, like, idx2: bool = False):
    """Length query -> caller-side allocation (where `like` lives) -> fetch; the handle is consumed."""
    k = lib.ab_trip_size(th)
    if k < 0:
        check(int(k))
    vv = empty_like_kind(like, (k,), np.float64)
    if idx2:
        ind = empty_like_kind(like, (k, 2), np.int64)
        check(lib.ab_trip_fetch_idx2(th, ptr(ind), ptr(vv)))
        return ind, vv
    ii = empty_like_kind(like, (k,), np.int64)
    jj = empty_like_kind(like, (k,), np.int64)
    check(lib.ab_trip_fetch(th, ptr(ii), ptr(jj), ptr(vv)))
    return ii, jj, vv


# ---------------------------------------------------------------------- op level (reference signatures)
def zero_out_row_op(indices, vv, bd):
    """ZeroOutRow(indices, vv, bd) -> (oindices, ovv)   (ZeroOutRow.cpp:14-19, ZeroOutRow.h:28-48)."""
    ki, pi = as_array(indices, np.int64)
    kv, pv = as_array(vv, np.float64)
    kb, pb = as_array(bd, np.int64)
    th = C.c_int64(0)
    check(lib.ab_zero_out_row(_n(kv), pi, pv, pb, _n(kb), C.byref(th)))
    return _fetch_trip(th.value, kv, idx2=True)


def zero_out_row_grad(grad_ovv, oindices, ovv, indices, vv, bd):
    """ZeroOutRowGrad -> (grad_indices, grad_vv, grad_bd); only grad_vv is defined (ZeroOutRow.h:50-57)."""
    kg, pg = as_array(grad_ovv, np.float64)
    ki, pi = as_array(indices, np.int64)
    kb, pb = as_array(bd, np.int64)
    N = _n(ki) // 2
    out = empty_like_kind(kg, (N,), np.float64)
    check(lib.ab_zero_out_row_grad(N, pi, pb, _n(kb), pg, _n(kg), ptr(out)))
    return None, out, None


def sparse_scatter_update_op(ii1, jj1, vv1, m1, n1, ii2, jj2, vv2, ii, jj):
    """SparseScatterUpdate -> (nii, njj, nvv)   (SparseScatterUpdate.cpp:20-34, .h:20-37)."""
    k1i, p1i = as_array(ii1, np.int64); k1j, p1j = as_array(jj1, np.int64); k1v, p1v = as_array(vv1, np.float64)
    k2i, p2i = as_array(ii2, np.int64); k2j, p2j = as_array(jj2, np.int64); k2v, p2v = as_array(vv2, np.float64)
    ki, pi = as_array(ii, np.int64); kj, pj = as_array(jj, np.int64)
    th = C.c_int64(0)
    check(lib.ab_scatter_update(_n(k1v), p1i, p1j, p1v, _n(k2v), p2i, p2j, p2v, int(m1), int(n1),
                                pi, _n(ki), pj, _n(kj), C.byref(th)))
    return _fetch_trip(th.value, k1v)


def sparse_scatter_update_grad(grad_nvv, nii, njj, nvv, ii1, jj1, vv1, m1, n1, ii2, jj2, vv2, ii, jj):
    """SparseScatterUpdateGrad -> 10 gradients, only grad_vv1 / grad_vv2 defined (.h:39-60)."""
    kg, pg = as_array(grad_nvv, np.float64)
    k1i, p1i = as_array(ii1, np.int64); k1j, p1j = as_array(jj1, np.int64)
    ki, pi = as_array(ii, np.int64); kj, pj = as_array(jj, np.int64)
    on, un = _n(k1i), _n(ii2)
    g1 = empty_like_kind(kg, (on,), np.float64)
    g2 = empty_like_kind(kg, (un,), np.float64)
    check(lib.ab_scatter_update_grad(on, p1i, p1j, un, int(m1), int(n1), pi, _n(ki), pj, _n(kj), pg, _n(kg),
                                     ptr(g1), ptr(g2)))
    return None, None, g1, None, None, None, None, g2, None, None


def _coo_handle(ii, jj, vv, m, n, base=1) -> int:
    ki, pi = as_array(ii, np.int64); kj, pj = as_array(jj, np.int64); kv, pv = as_array(vv, np.float64)
    h = C.c_int64(0)
    check(lib.ab_csr_from_coo(_n(kv), pi, pj, pv, int(m), int(n), base, C.byref(h)))
    return h.value


def sparse_indexing_op(ii1, jj1, vv1, m, n, ii, jj):
    """SparseIndexing -> (ii2, jj2, vv2)   (SparseIndexing.cpp:13-24, .h:23-63)."""
    h = _coo_handle(ii1, jj1, vv1, m, n)
    try:
        ki, pi = as_array(ii, np.int64); kj, pj = as_array(jj, np.int64)
        th = C.c_int64(0)
        check(lib.ab_sparse_indexing(h, pi, _n(ki), pj, _n(kj), C.byref(th)))
        like = vv1 if _is_torch(vv1) else np.empty(0)
        return _fetch_trip(th.value, like)
    finally:
        lib.ab_csr_destroy(h)


def sparse_indexing_grad(grad_vv2, ii2, jj2, vv2, ii1, jj1, vv1, m, n, ii, jj):
    """SparseIndexingGrad -> 7 gradients, only grad_vv1 defined (.h:66-83)."""
    h = _coo_handle(ii1, jj1, vv1, m, n)
    try:
        kg, pg = as_array(grad_vv2, np.float64)
        ki, pi = as_array(ii, np.int64); kj, pj = as_array(jj, np.int64)
        k0i, p0i = as_array(ii2, np.int64); k0j, p0j = as_array(jj2, np.int64)
        out = empty_like_kind(kg, (_n(ii1),), np.float64)
        check(lib.ab_sparse_indexing_grad(h, pi, _n(ki), pj, _n(kj), p0i, p0j, pg, _n(kg), ptr(out)))
        return None, None, out, None, None, None, None
    finally:
        lib.ab_csr_destroy(h)


def _matmul_colmajor(hA: int, hB: int, like):
    """(A*B) as 1-based triplets in column-major order — what the reference ops emit by walking the
    column-major Eigen result (SparseMatMul.cpp: `for k in outerSize: for InnerIterator`)."""
    hc = C.c_int64(0)
    check(lib.ab_spgemm(hA, hB, 1, C.byref(hc)))      # CSR of C' : its rows are C's columns
    try:
        m, n, nnz, T = (C.c_int64() for _ in range(4))
        check(lib.ab_csr_query(hc.value, C.byref(m), C.byref(n), C.byref(nnz), C.byref(T)))
        jj3 = empty_like_kind(like, (nnz.value,), np.int64)
        ii3 = empty_like_kind(like, (nnz.value,), np.int64)
        vv3 = empty_like_kind(like, (nnz.value,), np.float64)
        check(lib.ab_csr_export_coo(hc.value, ptr(jj3), ptr(ii3), ptr(vv3), 1))
        return ii3, jj3, vv3
    finally:
        lib.ab_csr_destroy(hc.value)


def sparse_sparse_mat_mul_op(ii1, jj1, vv1, ii2, jj2, vv2, m, n, k):
    """SparseSparseMatMul(ii1,jj1,vv1,ii2,jj2,vv2,m,n,k) -> (ii3,jj3,vv3); inputs 0-based, outputs
    1-based column-major (SparseMatMul.cpp:12-24; SparseMatMul.h:11-32)."""
    hA = _coo_handle(ii1, jj1, vv1, m, n, base=0)
    hB = _coo_handle(ii2, jj2, vv2, n, k, base=0)
    try:
        return _matmul_colmajor(hA, hB, vv1 if _is_torch(vv1) else np.empty(0))
    finally:
        lib.ab_csr_destroy(hA); lib.ab_csr_destroy(hB)


def _scaled_colmajor(ii, jj, vv, idx, d, m, k):
    kx, px = as_array(idx, np.int64); kv, pv = as_array(vv, np.float64); kd, pd = as_array(d, np.float64)
    out = empty_like_kind(kv, (_n(kv),), np.float64)
    check(lib.ab_coo_scale(_n(kv), px, pv, pd, _n(kd), 0, ptr(out)))
    # setFromTriplets of the scaled triplets, emitted column-major: assemble the transpose
    h = _coo_handle(jj, ii, out, k, m, base=0)
    try:
        nnz = C.c_int64()
        check(lib.ab_csr_query(h, None, None, C.byref(nnz), None))
        ii3, jj3 = empty_like_kind(kv, (nnz.value,), np.int64), empty_like_kind(kv, (nnz.value,), np.int64)
        vv3 = empty_like_kind(kv, (nnz.value,), np.float64)
        check(lib.ab_csr_export_coo(h, ptr(jj3), ptr(ii3), ptr(vv3), 1))
        return ii3, jj3, vv3
    finally:
        lib.ab_csr_destroy(h)


def diag_sparse_mat_mul_op(ii1, jj1, vv1, ii2, jj2, vv2, m, n, k):
    """DiagSparseMatMul: C = diag(vv1) * B, triplets (ii2, jj2, vv2*vv1[ii2]) (SparseMatMul.h:34-42)."""
    return _scaled_colmajor(ii2, jj2, vv2, ii2, vv1, m, k)


def sparse_diag_mat_mul_op(ii1, jj1, vv1, ii2, jj2, vv2, m, n, k):
    """SparseDiagMatMul: C = A * diag(vv2), triplets (ii1, jj1, vv1*vv2[jj1]) (SparseMatMul.h:44-52)."""
    return _scaled_colmajor(ii1, jj1, vv1, jj1, vv2, m, k)


# ---------------------------------------------------------------------- SparseTensor level
def _index_list(ix, dim: int):
    """Julia-style index argument -> 1-based int64 list (sparse.jl:305-314): Colon (None / ':' /
    slice(None)), integer, range / 1-based inclusive slice, integer array, Bool mask."""
    if ix is None or (isinstance(ix, str) and ix == ":") or (isinstance(ix, slice) and ix == slice(None)):
        return np.arange(1, dim + 1, dtype=np.int64), False
    if isinstance(ix, (int, np.integer)) and not isinstance(ix, (bool, np.bool_)):
        return np.array([int(ix)], dtype=np.int64), True
    if isinstance(ix, slice):       # a:b and a:s:b are Julia ranges: 1-based, inclusive
        step = 1 if ix.step is None else int(ix.step)
        lo = 1 if ix.start is None else int(ix.start)
        hi = dim if ix.stop is None else int(ix.stop)
        return np.arange(lo, hi + (1 if step > 0 else -1), step, dtype=np.int64), False
    if isinstance(ix, range):
        return np.asarray(list(ix), dtype=np.int64), False
    a = ix.cpu().numpy() if _is_torch(ix) else np.asarray(ix)
    if a.dtype == np.bool_:
        return (np.flatnonzero(a) + 1).astype(np.int64), False     # findall
    return a.astype(np.int64).ravel(), False


def _autograd():
    import torch

    class ZeroOutRowFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, vv, indices, bd):
            oind, ovv = zero_out_row_op(indices, vv.detach(), bd)
            ctx.indices, ctx.bd = indices, bd
            ctx.mark_non_differentiable(oind)
            return oind, ovv

        @staticmethod
        def backward(ctx, _gi, govv):
            return zero_out_row_grad(govv.contiguous(), None, None, ctx.indices, None, ctx.bd)[1], None, None

    class ScatterUpdateFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, vv1, vv2, ii1, jj1, m, n, ii2, jj2, ii, jj):
            nii, njj, nvv = sparse_scatter_update_op(ii1, jj1, vv1.detach(), m, n, ii2, jj2, vv2.detach(), ii, jj)
            ctx.args = (ii1, jj1, m, n, ii2, jj2, ii, jj)
            ctx.mark_non_differentiable(nii, njj)
            return nii, njj, nvv

        @staticmethod
        def backward(ctx, _a, _b, g):
            ii1, jj1, m, n, ii2, jj2, ii, jj = ctx.args
            r = sparse_scatter_update_grad(g.contiguous(), None, None, None, ii1, jj1, None, m, n, ii2, jj2, None, ii, jj)
            return (r[2], r[7]) + (None,) * 8

    class IndexingFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, vv1, ii1, jj1, m, n, ii, jj):
            ii2, jj2, vv2 = sparse_indexing_op(ii1, jj1, vv1.detach(), m, n, ii, jj)
            ctx.args = (ii1, jj1, m, n, ii, jj, ii2, jj2)
            ctx.save_for_backward(vv1)
            ctx.mark_non_differentiable(ii2, jj2)
            return ii2, jj2, vv2

        @staticmethod
        def backward(ctx, _a, _b, g):
            ii1, jj1, m, n, ii, jj, ii2, jj2 = ctx.args
            (vv1,) = ctx.saved_tensors
            r = sparse_indexing_grad(g.contiguous(), ii2, jj2, None, ii1, jj1, vv1.detach(), m, n, ii, jj)
            return (r[2],) + (None,) * 6

    return ZeroOutRowFn, ScatterUpdateFn, IndexingFn


_AG = None


def _ag():
    global _AG
    if _AG is None:
        _AG = _autograd()
    return _AG


def _req(x) -> bool:
    return _is_torch(x) and x.requires_grad


def _sparse():
    from . import sparse as sp

    return sp


def _as_st(A):
    sp = _sparse()
    return A if isinstance(A, sp.SparseTensor) else sp.SparseTensor(A)


def zero_out_row(A, bd):
    """zero_out_row(A, bd) (sparse.jl:790-797): drop every stored entry in the rows `bd` (1-based)."""
    sp = _sparse()
    A = _as_st(A)
    ii, jj, vv = A.triplets()
    if _is_torch(vv):
        import torch

        ind = torch.stack([torch.as_tensor(ii, device=vv.device).to(torch.int64) - 1,
                           torch.as_tensor(jj, device=vv.device).to(torch.int64) - 1], dim=1).contiguous()
    else:
        ind = np.stack([np.asarray(ii, np.int64) - 1, np.asarray(jj, np.int64) - 1], axis=1)
    bd = np.asarray(bd, dtype=np.int64).ravel()
    if _req(vv):
        oind, ovv = _ag()[0].apply(vv, ind, bd)
    else:
        oind, ovv = zero_out_row_op(ind, vv, bd)
    return sp.SparseTensor(oind[:, 0] + 1, oind[:, 1] + 1, ovv, A.m, A.n)


def scatter_update(A, i1, i2, B):
    """A[i1, i2] = B (sparse.jl:340-364)."""
    sp = _sparse()
    A, B = _as_st(A), _as_st(B)
    ii, _ = _index_list(i1, A.m)
    jj, _ = _index_list(i2, A.n)
    ii1, jj1, vv1 = A.triplets()
    ii2, jj2, vv2 = B.triplets()
    if _req(vv1) or _req(vv2):
        import torch

        dev = vv1.device if _is_torch(vv1) else vv2.device
        vv1 = vv1 if _is_torch(vv1) else torch.as_tensor(np.asarray(vv1, np.float64), device=dev)
        vv2 = vv2 if _is_torch(vv2) else torch.as_tensor(np.asarray(vv2, np.float64), device=dev)
        nii, njj, nvv = _ag()[1].apply(vv1, vv2, ii1, jj1, A.m, A.n, ii2, jj2, ii, jj)
    else:
        nii, njj, nvv = sparse_scatter_update_op(ii1, jj1, vv1, A.m, A.n, ii2, jj2, vv2, ii, jj)
    return sp.SparseTensor(nii, njj, nvv, A.m, A.n)


def scatter_add(A, i1, i2, B):
    """A[i1, i2] += B (sparse.jl:377-386): C = A[i1,i2]; D = B + C; scatter_update(A, i1, i2, D)."""
    A, B = _as_st(A), _as_st(B)
    return scatter_update(A, i1, i2, add(B, getindex(A, i1, i2)))


def getindex(A, i1, i2):
    """A[i1, i2] (sparse.jl:303-327).  Integer arguments squeeze that dimension and the result is dense
    (the reference returns squeeze(Array(ret)))."""
    sp = _sparse()
    A = _as_st(A)
    ii, sq1 = _index_list(i1, A.m)
    jj, sq2 = _index_list(i2, A.n)
    ii1, jj1, vv1 = A.triplets()
    if _req(vv1):
        ii2, jj2, vv2 = _ag()[2].apply(vv1, ii1, jj1, A.m, A.n, ii, jj)
    else:
        ii2, jj2, vv2 = sparse_indexing_op(ii1, jj1, vv1, A.m, A.n, ii, jj)
    ret = sp.SparseTensor(ii2, jj2, vv2, len(ii), len(jj))
    if sq1 or sq2:
        d = ret.toarray()
        return d.squeeze(axis=tuple(k for k, s in enumerate((sq1, sq2)) if s))
    return ret


def _cat(a, b, dtype):
    if _is_torch(a) or _is_torch(b):
        import torch

        dev = a.device if _is_torch(a) else b.device
        tdt = torch.float64 if dtype == np.float64 else torch.int64
        ta = a if _is_torch(a) else torch.as_tensor(np.asarray(a, dtype), device=dev)
        tb = b if _is_torch(b) else torch.as_tensor(np.asarray(b, dtype), device=dev)
        return torch.cat([ta.to(tdt), tb.to(tdt)])
    return np.concatenate([np.asarray(a, dtype), np.asarray(b, dtype)])


def add(A, B, alpha: float = 1.0, beta: float = 1.0):
    """A + B on the union pattern (sparse.jl:217 -> tf.sparse.add).  With triplets at hand the sum is
    the concatenated stream (duplicates are summed by every consumer, A's term first); two
    pattern-only (device-handle) operands go through ab_csr_axpby."""
    sp = _sparse()
    A, B = _as_st(A), _as_st(B)
    if A.shape != B.shape:
        raise ValueError(f"size {A.shape} and {B.shape} does not match")
    if A._pattern_only and B._pattern_only:
        h = C.c_int64(0)
        check(lib.ab_csr_axpby(A.handle(), float(alpha), B.handle(), float(beta), C.byref(h)))
        return sp.SparseTensor.from_handle(h.value)
    ii1, jj1, vv1 = A.triplets()
    ii2, jj2, vv2 = B.triplets()
    v1 = vv1 if alpha == 1.0 else vv1 * alpha
    v2 = vv2 if beta == 1.0 else vv2 * beta
    return sp.SparseTensor(_cat(ii1, ii2, np.int64), _cat(jj1, jj2, np.int64), _cat(v1, v2, np.float64), A.m, A.n,
                           is_diag=A._diag and B._diag)


def matmul(A, B):
    """SparseTensor * SparseTensor (sparse.jl:579-595): SpGEMM, with the reference's diagonal fast
    paths when either factor carries `_diag`.  No gradient is registered (as in the reference)."""
    sp = _sparse()
    A, B = _as_st(A), _as_st(B)
    if A.n != B.m:
        raise ValueError(f"matrix size mismatch: ({A.m}, {A.n}) vs ({B.m}, {B.n})")
    if A._diag and not A._pattern_only and not B._pattern_only and _n(A.vv) == A.m:
        ii2, jj2, vv2 = B.triplets()
        i3, j3, v3 = diag_sparse_mat_mul_op(None, None, _diag_values(A), _z(ii2), _z(jj2), vv2, A.m, A.n, B.n)
        return sp.SparseTensor(i3, j3, v3, A.m, B.n, is_diag=A._diag and B._diag)
    if B._diag and not A._pattern_only and not B._pattern_only and _n(B.vv) == B.m:
        ii1, jj1, vv1 = A.triplets()
        i3, j3, v3 = sparse_diag_mat_mul_op(_z(ii1), _z(jj1), vv1, None, None, _diag_values(B), A.m, A.n, B.n)
        return sp.SparseTensor(i3, j3, v3, A.m, B.n, is_diag=False)
    h = C.c_int64(0)
    check(lib.ab_spgemm(A.handle(), B.handle(), 0, C.byref(h)))
    return sp.SparseTensor.from_handle(h.value)


def _z(ix):
    """1-based -> 0-based index array (the matmul ops take `ii-1`, sparse.jl:593)."""
    return ix - 1 if _is_torch(ix) else np.asarray(ix, np.int64) - 1


def _diag_values(D):
    """Values of a `_diag` SparseTensor ordered by row — the reference indexes vv by row directly,
    which assumes this order (spdiag builds it so, sparse.jl:553-559)."""
    ii, jj, vv = D.triplets()
    if _is_torch(vv):
        import torch

        out = torch.zeros(D.m, dtype=torch.float64, device=vv.device)
        out[torch.as_tensor(ii, device=vv.device).to(torch.int64) - 1] = vv.detach()
        return out
    out = np.zeros(D.m)
    out[np.asarray(ii, np.int64) - 1] = np.asarray(vv, np.float64)
    return out


def sparse_concate_op(ii1, jj1, vv1, m1, n1, ii2, jj2, vv2, m2, n2, hcat_):
    """SparseConcate(ii1,jj1,vv1,m1,n1,ii2,jj2,vv2,m2,n2,hcat) -> (ii,jj,vv)  (SparseConcate.cpp:19-34,
    SparseConcate.h:11-27): the second operand's rows (vcat) or columns (hcat) are shifted, order kept."""
    k1i, p1i = as_array(ii1, np.int64); k1j, p1j = as_array(jj1, np.int64); k1v, p1v = as_array(vv1, np.float64)
    k2i, p2i = as_array(ii2, np.int64); k2j, p2j = as_array(jj2, np.int64); k2v, p2v = as_array(vv2, np.float64)
    n = _n(k1v) + _n(k2v)
    like = k1v if _is_torch(k1v) else k2v
    oi, oj, ov = empty_like_kind(like, (n,), np.int64), empty_like_kind(like, (n,), np.int64), empty_like_kind(like, (n,), np.float64)
    check(lib.ab_coo_concat(_n(k1v), p1i, p1j, p1v, _n(k2v), p2i, p2j, p2v, int(m1), int(n1), 1 if hcat_ else 0,
                            ptr(oi), ptr(oj), ptr(ov)))
    return oi, oj, ov


def sparse_concate_grad(grad_vv, ii, jj, vv, ii1, jj1, vv1, m1, n1, ii2, jj2, vv2, m2, n2, hcat_):
    """SparseConcateGrad: the gradient is the split of grad_vv (SparseConcate.h:60-68)."""
    N1 = _n(ii1)
    return None, None, grad_vv[:N1], None, None, None, None, grad_vv[N1:], None, None, None


def _concat(A, B, hcat_: bool):
    sp = _sparse()
    A, B = _as_st(A), _as_st(B)
    if hcat_ and A.m != B.m or not hcat_ and A.n != B.n:
        raise ValueError(f"sparse_concate: shapes {A.shape} and {B.shape} do not fit")
    ii1, jj1, vv1 = A.triplets()
    ii2, jj2, vv2 = B.triplets()
    if _req(vv1) or _req(vv2):                 # the split gradient is torch.cat's own
        dI, dJ = (0, A.n) if hcat_ else (A.m, 0)
        ii = _cat(ii1, (ii2 + dI), np.int64); jj = _cat(jj1, (jj2 + dJ), np.int64); vv = _cat(vv1, vv2, np.float64)
    else:
        ii, jj, vv = sparse_concate_op(ii1, jj1, vv1, A.m, A.n, ii2, jj2, vv2, B.m, B.n, hcat_)
    return sp.SparseTensor(ii, jj, vv, A.m if hcat_ else A.m + B.m, A.n + B.n if hcat_ else A.n)


def hcat(*args):
    """[A B ...] (sparse.jl:247-249,258)."""
    out = args[0]
    for B in args[1:]:
        out = _concat(out, B, True)
    return out


def vcat(*args):
    """[A; B; ...] (sparse.jl:247-249,257)."""
    out = args[0]
    for B in args[1:]:
        out = _concat(out, B, False)
    return out


def hvcat(rows, *values):
    """[A B; C] (sparse.jl:259-268): `rows` gives the number of blocks in each block row."""
    r, k = [], 0
    for cnt in rows:
        r.append(hcat(*values[k:k + cnt]))
        k += cnt
    return vcat(*r)


def dense_to_sparse(o):
    """dense_to_sparse(o) (sparse.jl:78-86): SparseTensor of the nonzeros of a dense matrix."""
    sp = _sparse()
    if len(tuple(o.shape)) != 2:
        raise ValueError("dense_to_sparse: a matrix is required")
    m, n = int(o.shape[0]), int(o.shape[1])
    ko, po = as_array(o, np.float64)
    th = C.c_int64(0)
    check(lib.ab_dense_to_coo(m, n, po, C.byref(th)))
    ii, jj, vv = _fetch_trip(th.value, ko if _is_torch(ko) else np.empty(0))
    return sp.SparseTensor(ii, jj, vv, m, n)


def to_dense(A):
    """Array(A) (sparse.jl:185-193 -> sparse_to_dense_ad): dense row-major matrix, duplicates summed."""
    A = _as_st(A)
    like = A.vv if (A.vv is not None and _is_torch(A.vv)) else np.empty(0)
    out = empty_like_kind(like, (A.m, A.n), np.float64)
    check(lib.ab_csr_to_dense(A.handle(), ptr(out)))
    return out


def sparse_to_dense_grad(grad_A, ij, m, n):
    """SparseToDenseAdGrad: grad_vv[i] = grad_A[ij[i,0], ij[i,1]] (SparseToDense.h:8-18)."""
    kg, pg = as_array(grad_A, np.float64)
    ki, pi = as_array(ij, np.int64)
    T = _n(ki) // 2
    out = empty_like_kind(kg, (T,), np.float64)
    check(lib.ab_dense_grad_values(T, pi, int(m), int(n), pg, ptr(out)))
    return out


def spsum(A, dims=None):
    """sum(S; dims) (sparse.jl:606-612 -> tf.sparse.reduce_sum): total, or the column sums (dims=1) /
    row sums (dims=2) as a vector — SpMV with a vector of ones."""
    sp = _sparse()
    A = _as_st(A)
    if dims is None:
        return float(np.asarray(sp.mul(A, np.ones(A.n))).sum())
    if dims == 1:
        return sp.rmul(np.ones(A.m), A)
    if dims == 2:
        return sp.mul(A, np.ones(A.n))
    raise ValueError("dims must be 1, 2 or None")


# ---------------------------------------------------------------------- least squares  A \ f,  A tall
def _normal_equations(A):
    """(handle of A' , SparseTensor of M = A'A) on the device: ab_csr_transpose + ab_spgemm."""
    sp = _sparse()
    ht = C.c_int64(0)
    check(lib.ab_csr_transpose(A.handle(), C.byref(ht)))
    At = sp.SparseTensor.from_handle(ht.value)
    hm = C.c_int64(0)
    check(lib.ab_spgemm(At.handle(), A.handle(), 0, C.byref(hm)))
    return At, sp.SparseTensor.from_handle(hm.value)


def _axpby(a, x, b, y):
    kx, px = as_array(x, np.float64); ky, py = as_array(y, np.float64)
    out = empty_like_kind(kx, (_n(kx),), np.float64)
    check(lib.ab_vec_axpby(_n(kx), float(a), px, float(b), py, ptr(out)))
    return out


def least_square_raw(A, f):
    """u = argmin ||A u - f||_2 for every row of f (SparseLeastSquare.h:10-32; the reference factors A with
    SparseQR).  Here: normal equations M = A'A (SpGEMM on the device), rhs = A'f, Jacobi-PCG on the SPD M —
    A must have full column rank, as SparseQR's solution being unique requires."""
    sp = _sparse()
    At, M = _normal_equations(A)
    vec = len(tuple(f.shape)) == 1
    rows = [f] if vec else [f[i] for i in range(int(f.shape[0]))]
    outs = []
    for fi in rows:
        if int(fi.shape[0]) != A.m:
            raise ValueError("dimension mismatch")
        rhs = sp._spmm_raw(At, fi if not _is_torch(fi) else fi.contiguous(), False)
        outs.append(sp._solve_raw(M, rhs, "CG"))
    if vec:
        return outs[0]
    if _is_torch(outs[0]):
        import torch

        return torch.stack(outs, dim=0)
    return np.stack(outs, axis=0)


def least_square_grad(A, grad_u, u, f):
    """SparseLeastSquare backward (SparseLeastSquare.h:34-83), one batch row: x = (A'A)^-1 g ; grad_f = A x ;
    grad_vv[k] = (f - A u)[i_k] x[j_k] - u[j_k] (A x)[i_k]  — two per-triplet outer-product gathers."""
    sp = _sparse()
    At, M = _normal_equations(A)
    x = sp._solve_raw(M, grad_u, "CG")
    gf = sp._spmm_raw(A, x, False)
    t = sp._spmm_raw(A, u, False)
    resid = _axpby(1.0, f, -1.0, t)
    T = A.query()[3]
    kx, px = as_array(x, np.float64); ku, pu = as_array(u, np.float64)
    kr, pr = as_array(resid, np.float64); kg, pg = as_array(gf, np.float64)
    g1 = empty_like_kind(kx, (T,), np.float64); g2 = empty_like_kind(kx, (T,), np.float64)
    check(lib.ab_spmm_grad_values(A.handle(), pr, px, 1, ptr(g1)))        # (f - A u)[i_k] * x[j_k]
    check(lib.ab_spmm_grad_values(A.handle(), pg, pu, 1, ptr(g2)))        # (A x)[i_k] * u[j_k]
    return _axpby(1.0, g1, -1.0, g2), gf


def least_square_autograd(A, f):
    """Differentiable least-squares `\\` (gradients for the values and the right-hand side)."""
    import torch

    dev = A.vv.device if _is_torch(A.vv) else f.device
    vv = A.vv if _is_torch(A.vv) else torch.as_tensor(np.asarray(A.vv, np.float64), device=dev)
    if not _is_torch(f):
        f = torch.as_tensor(np.asarray(f, np.float64), device=dev)

    class LeastSquareFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, vv_, f_):
            u = least_square_raw(A, f_.detach())
            ctx.save_for_backward(u, f_)
            return u

        @staticmethod
        def backward(ctx, gu):
            u, f_ = ctx.saved_tensors
            gu = gu.contiguous()
            if f_.dim() == 1:
                gvv, gf = least_square_grad(A, gu, u, f_.detach())
                return gvv, gf
            gvv, gfs = None, []
            for l in range(f_.shape[0]):
                g1, gf = least_square_grad(A, gu[l].contiguous(), u[l].contiguous(), f_[l].detach().contiguous())
                gvv = g1 if gvv is None else gvv + g1
                gfs.append(gf)
            return gvv, torch.stack(gfs, 0)

    return LeastSquareFn.apply(vv, f)


def sparse_least_square_op(ii, jj, vv, ff, n):
    """SparseLeastSquare(ii, jj, vv, ff, n) -> u ; ff is [nbatch, m] (SparseLeastSquare.cpp:11-18)."""
    sp = _sparse()
    m = int(ff.shape[-1])
    A = sp.SparseTensor(ii, jj, vv, m, int(n))
    return least_square_raw(A, ff)


def sparse_least_square_grad(grad_u, u, ii, jj, vv, ff, n):
    """SparseLeastSquareGrad -> (grad_ii, grad_jj, grad_vv, grad_ff, grad_n); batch rows accumulate into grad_vv."""
    sp = _sparse()
    m = int(ff.shape[-1])
    A = sp.SparseTensor(ii, jj, vv, m, int(n))
    if len(tuple(ff.shape)) == 1:
        gvv, gf = least_square_grad(A, grad_u, u, ff)
        return None, None, gvv, gf, None
    gvv, gfs = None, []
    for l in range(int(ff.shape[0])):
        g1, gf = least_square_grad(A, grad_u[l], u[l], ff[l])
        gvv = g1 if gvv is None else _axpby(1.0, gvv, 1.0, g1)
        gfs.append(gf)
    if _is_torch(gfs[0]):
        import torch

        return None, None, gvv, torch.stack(gfs, 0), None
    return None, None, gvv, np.stack(gfs, 0), None


def tri_solve_op(a, b, c, d):
    """TriSolve(a, b, c, d) -> x   (TriSolve.cpp:14-19, TriSolve.h:1-20)."""
    ka, pa = as_array(a, np.float64); kb, pb = as_array(b, np.float64)
    kc, pc = as_array(c, np.float64); kd, pd = as_array(d, np.float64)
    n = _n(kb)
    x = empty_like_kind(kd, (n,), np.float64)
    check(lib.ab_tri_solve(n, pa, pb, pc, pd, ptr(x)))
    return x


def tri_solve_grad(grad_x, x, a, b, c, d):
    """TriSolveGrad(grad_x, x, a, b, c, d) -> (grad_a, grad_b, grad_c, grad_d)   (TriSolve.cpp:35-45, .h:22-37)."""
    kg, pg = as_array(grad_x, np.float64); kx, px = as_array(x, np.float64)
    ka, pa = as_array(a, np.float64); kb, pb = as_array(b, np.float64); kc, pc = as_array(c, np.float64)
    n = _n(kb)
    ga, gc = empty_like_kind(kg, (max(n - 1, 0),), np.float64), empty_like_kind(kg, (max(n - 1, 0),), np.float64)
    gb, gd = empty_like_kind(kg, (n,), np.float64), empty_like_kind(kg, (n,), np.float64)
    check(lib.ab_tri_solve_grad(n, pg, px, pa, pb, pc, ptr(ga), ptr(gb), ptr(gc), ptr(gd)))
    return ga, gb, gc, gd


def trisolve(a, b, c, d):
    """trisolve(a, b, c, d) (sparse.jl:731-739): solve the tridiagonal system with sub-diagonal a, diagonal b,
    super-diagonal c and right-hand side d."""
    n = _n(b)
    if not (_n(d) == n and _n(a) == n - 1 and _n(c) == n - 1):
        raise ValueError("trisolve: need length(b) == length(d) == length(a)+1 == length(c)+1")
    if any(_req(t) for t in (a, b, c, d)):
        import torch

        dev = next(t.device for t in (a, b, c, d) if _is_torch(t))
        ts = [t if _is_torch(t) else torch.as_tensor(np.asarray(t, np.float64), device=dev) for t in (a, b, c, d)]

        class TriSolveFn(torch.autograd.Function):
            @staticmethod
            def forward(ctx, a_, b_, c_, d_):
                x = tri_solve_op(a_.detach(), b_.detach(), c_.detach(), d_.detach())
                ctx.save_for_backward(x, a_, b_, c_)
                return x

            @staticmethod
            def backward(ctx, gx):
                x, a_, b_, c_ = ctx.saved_tensors
                return tri_solve_grad(gx.contiguous(), x, a_.detach(), b_.detach(), c_.detach(), None)

        return TriSolveFn.apply(*ts)
    return tri_solve_op(a, b, c, d)


def spdiag(*args):
    """spdiag(n) identity, spdiag(o) diagonal matrix from a vector, spdiag(n, k1=>v1, ...) banded
    matrix from (offset, values) pairs (sparse.jl:544-561, 631-658)."""
    sp = _sparse()
    if len(args) == 1 and isinstance(args[0], (int, np.integer)):
        n = int(args[0])
        i = np.arange(1, n + 1, dtype=np.int64)
        return sp.SparseTensor(i, i.copy(), np.ones(n), n, n)
    if len(args) == 1:
        o = args[0]
        if len(tuple(o.shape)) != 1:
            raise ValueError("ADCME: input `o` must be a vector")
        n = int(o.shape[0])
        i = np.arange(1, n + 1, dtype=np.int64)
        return sp.SparseTensor(i, i.copy(), o, n, n, is_diag=True)
    n, pairs = int(args[0]), args[1:]
    I, J, V = [], [], []
    for k, v in pairs:
        k = int(k)
        if not -(n - 1) <= k <= n - 1:
            raise ValueError("spdiag: offset out of range")
        if _n(v) != n - abs(k):
            raise ValueError(f"spdiag: diagonal {k} needs {n - abs(k)} values")
        if k >= 0:
            I.append(np.arange(1, n - k + 1, dtype=np.int64)); J.append(np.arange(k + 1, n + 1, dtype=np.int64))
        else:
            I.append(np.arange(-k + 1, n + 1, dtype=np.int64)); J.append(np.arange(1, n + k + 1, dtype=np.int64))
        V.append(v)
    if any(_is_torch(v) for v in V):
        import torch

        dev = next(v.device for v in V if _is_torch(v))
        vv = torch.cat([v if _is_torch(v) else torch.as_tensor(np.asarray(v, np.float64), device=dev) for v in V])
    else:
        vv = np.concatenate([np.asarray(v, np.float64) for v in V])
    return sp.SparseTensor(np.concatenate(I), np.concatenate(J), vv, n, n)


def spzero(m: int, n: int | None = None):
    """spzero(m, n=m) (sparse.jl:568-575)."""
    sp = _sparse()
    n = m if n is None else n
    return sp.SparseTensor(np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0), int(m), int(n), is_diag=True)
