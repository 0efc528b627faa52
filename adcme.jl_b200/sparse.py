"""Host-side mirror of /root/reference/src/sparse.jl for the hot path, over the C-ABI.

Julia is not available in the build image, so this Python module plays the role the reference's
Julia layer plays (same names, argument meaning and error behaviour); julia/ADCMEB200.jl holds the
`ccall` shim a maintainer would drop into ADCME itself.  Nothing here computes: every operation
is one call into libadcme_b200.so.  torch is used for device buffers and, when values require
grad, as the host AD engine that invokes the `*_grad` entry points — the role TensorFlow's
`tf.custom_gradient` plays in the reference (src/extra.jl:160-235).

    SparseTensor(I, J, V, m, n)      src/sparse.jl:62-76     (1-based I, J)
    find / rows / cols / values      src/sparse.jl:37-47,93-99
    A * x , x * A , A'               src/sparse.jl:220-264
    A \\ f  (backslash / solve)       src/sparse.jl:413-445
    SparseAssembler/accumulate/assemble   src/sparse.jl:480-534
    compress                         src/sparse.jl:776-781
    factorize / solve                src/sparse.jl:671-699
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import ADCMEError, as_array, check, empty_like_kind, lib, ptr
from .options import options


def _is_torch(x) -> bool:
    return _lib._is_torch(x)


def _numel(x) -> int:
    return int(x.numel()) if _is_torch(x) else int(np.asarray(x).size)


def _shape(x):
    return tuple(x.shape)


class SparseTensor:
    """COO triplets (1-based, duplicates allowed) + a lazily assembled device CSR handle."""

    __array_ufunc__ = None      # numpy defers `ndarray @ SparseTensor` to __rmatmul__
    __array_priority__ = 1000

    def __init__(self, I, J=None, V=None, m=None, n=None, is_diag: bool = False):
        if J is None and hasattr(I, "tocoo"):          # SparseTensor(A::SparseMatrixCSC), sparse.jl:144-157
            coo = I.tocoo()
            m, n = coo.shape
            I, J, V = coo.row.astype(np.int64) + 1, coo.col.astype(np.int64) + 1, coo.data.astype(np.float64)
        if J is None and V is None and len(tuple(getattr(I, "shape", ()))) == 2:   # SparseTensor(A::Array), sparse.jl:159-161
            from . import structure

            d = structure.dense_to_sparse(I)
            I, J, V, m, n = d.ii, d.jj, d.vv, d.m, d.n
        if _numel(I) != _numel(J) or _numel(I) != _numel(V):
            raise ValueError("I, J, V must have the same length")
        self.ii, self.jj, self.vv = I, J, V
        if m is None:
            m = int(np.max(np.asarray(I))) if _numel(I) else 0
        if n is None:
            n = int(np.max(np.asarray(J))) if _numel(J) else 0
        self.m, self.n = int(m), int(n)
        self._diag = bool(is_diag)
        self._h = None          # CSR handle
        self._pattern_only = False

    # ------------------------------------------------------------------ handle management
    @classmethod
    def from_handle(cls, h: int):
        """Wrap an existing CSR handle (e.g. from a stencil generator); triplets == stored entries."""
        self = cls.__new__(cls)
        m, n, nnz, T = (C.c_int64() for _ in range(4))
        check(lib.ab_csr_query(h, C.byref(m), C.byref(n), C.byref(nnz), C.byref(T)))
        self.ii = self.jj = self.vv = None
        self.m, self.n, self._diag, self._h, self._pattern_only = m.value, n.value, False, int(h), True
        return self

    def handle(self) -> int:
        if self._h is None:
            ki, pi = as_array(self.ii, np.int64)
            kj, pj = as_array(self.jj, np.int64)
            kv, pv = as_array(self.vv, np.float64)
            h = C.c_int64(0)
            check(lib.ab_csr_from_coo(_numel(self.vv), pi, pj, pv, self.m, self.n, 1, C.byref(h)))
            self._h = h.value
        return self._h

    def _refresh_values(self, vv) -> None:
        """New values on the cached pattern: re-reduce through the stored map, no re-sort."""
        h = self.handle()
        kv, pv = as_array(vv, np.float64)
        check(lib.ab_csr_update_values(h, pv))

    def close(self) -> None:
        if self._h is not None:
            try:
                lib.ab_csr_destroy(self._h)
            finally:
                self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ introspection
    @property
    def shape(self):
        return (self.m, self.n)

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim - 1]   # 1-based like Julia's size(s, i)

    def query(self):
        m, n, nnz, T = (C.c_int64() for _ in range(4))
        check(lib.ab_csr_query(self.handle(), C.byref(m), C.byref(n), C.byref(nnz), C.byref(T)))
        return m.value, n.value, nnz.value, T.value

    def csr(self):
        """(rowptr, colind, values) of the assembled matrix as numpy arrays (duplicates summed)."""
        m, n, nnz, T = self.query()
        rp, ci, va = np.empty(m + 1, np.int32), np.empty(nnz, np.int32), np.empty(nnz, np.float64)
        check(lib.ab_csr_export(self.handle(), ptr(rp), ptr(ci), ptr(va)))
        return rp, ci, va

    def triplets(self):
        """(ii, jj, vv), 1-based, in stored order.  A pattern-only tensor (built from a device handle)
        materialises them once from its CSR (row-major, duplicates already summed)."""
        if self.ii is None:
            m, n, nnz, T = self.query()
            ii, jj, vv = np.empty(nnz, np.int64), np.empty(nnz, np.int64), np.empty(nnz, np.float64)
            check(lib.ab_csr_export_coo(self.handle(), ptr(ii), ptr(jj), ptr(vv), 1))
            self.ii, self.jj, self.vv = ii, jj, vv
        return self.ii, self.jj, self.vv

    def slot_map(self):
        m, n, nnz, T = self.query()
        s = np.empty(T, np.int32)
        check(lib.ab_csr_export_map(self.handle(), ptr(s)))
        return s

    def to_scipy(self):
        """run(sess, S) (sparse.jl:175-178): the matrix with duplicates summed."""
        import scipy.sparse as sp

        rp, ci, va = self.csr()
        return sp.csr_matrix((va, ci, rp), shape=self.shape)

    def toarray(self):
        return self.to_scipy().toarray()

    # ------------------------------------------------------------------ algebra
    def adjoint(self):
        """A' = swapped indices (sparse.jl:220-225)."""
        if self._pattern_only:
            h = C.c_int64(0)
            check(lib.ab_csr_transpose(self.handle(), C.byref(h)))
            return SparseTensor.from_handle(h.value)
        return SparseTensor(self.jj, self.ii, self.vv, self.n, self.m, is_diag=self._diag)

    T = property(adjoint)

    def __matmul__(self, o):
        if isinstance(o, SparseTensor):
            from . import structure

            return structure.matmul(self, o)
        return mul(self, o)

    def __mul__(self, o):
        if np.isscalar(o):
            ii, jj, vv = self.triplets()
            return SparseTensor(ii, jj, vv * float(o), self.m, self.n, is_diag=self._diag)
        return self.__matmul__(o)

    def __truediv__(self, o):
        return self * (1.0 / float(o))          # sparse.jl:225

    def __add__(self, o):
        from . import structure

        if isinstance(o, SparseTensor) or hasattr(o, "tocoo"):
            return structure.add(self, o)       # sparse.jl:217
        if tuple(o.shape) != self.shape:        # sparse.jl:203-209: SparseTensor + dense -> dense
            raise ValueError(f"size {self.shape} and {tuple(o.shape)} does not match")
        return self.toarray() + (o.detach().cpu().numpy() if _is_torch(o) else np.asarray(o))

    __radd__ = __add__

    def __sub__(self, o):
        if isinstance(o, SparseTensor) or hasattr(o, "tocoo"):
            from . import structure

            return structure.add(self, o, 1.0, -1.0)   # s1 + (-s2), sparse.jl:218
        return self + (-o)

    def __rsub__(self, o):
        return (-self) + o

    def __getitem__(self, key):
        from . import structure

        if not isinstance(key, tuple) or len(key) != 2:
            raise IndexError("SparseTensor indexing takes two indices: A[i1, i2]")
        return structure.getindex(self, key[0], key[1])

    def __rmatmul__(self, o):
        return rmul(o, self)

    def __rmul__(self, o):
        if np.isscalar(o):
            return self * o
        return rmul(o, self)

    def __neg__(self):
        return self * -1.0

    def solve(self, f, method: str = "auto"):
        return backslash(self, f, method)


# ---------------------------------------------------------------------- find / rows / cols / values
def find(s: SparseTensor):
    """(ii, jj, vv), 1-based; row-major sorted when options.sparse.auto_reorder (tf.sparse.reorder
    is a stable lexicographic sort, duplicates kept: sparse.jl:62-76,93-99)."""
    ii, jj, vv = s.triplets()
    if options.sparse.auto_reorder and not s._pattern_only:      # a handle's triplets are already row-major
        ki, pi = as_array(ii, np.int64)
        kj, pj = as_array(jj, np.int64)
        T = _numel(ki)
        like = vv if _is_torch(vv) else (ki if _is_torch(ki) else np.empty(0))
        oi, oj = empty_like_kind(like, (T,), np.int64), empty_like_kind(like, (T,), np.int64)
        perm = empty_like_kind(like, (T,), np.int64)
        if _is_torch(vv) and vv.requires_grad:                   # keep the autograd edge: permute on the AD side
            check(lib.ab_coo_reorder(T, pi, pj, None, s.m, s.n, 1, ptr(oi), ptr(oj), None, ptr(perm)))
            vv = vv[perm]
        else:
            kv, pv = as_array(vv, np.float64)
            ov = empty_like_kind(like, (T,), np.float64)
            check(lib.ab_coo_reorder(T, pi, pj, pv, s.m, s.n, 1, ptr(oi), ptr(oj), ptr(ov), ptr(perm)))
            vv = ov
        ii, jj = oi, oj
    return ii, jj, vv


def rows(s):
    return find(s)[0]


def cols(s):
    return find(s)[1]


def values(s):
    return find(s)[2]


# ---------------------------------------------------------------------- A * x
def _spmm_raw(A: SparseTensor, X, trans: bool):
    vec = len(_shape(X)) == 1
    k = 1 if vec else int(_shape(X)[1])
    out_rows = A.n if trans else A.m
    in_rows = A.m if trans else A.n
    if int(_shape(X)[0]) != in_rows:
        raise ValueError(f"dimension mismatch: matrix is {A.m}x{A.n}, operand has {_shape(X)[0]} rows")
    kx, px = as_array(X, np.float64)
    Y = empty_like_kind(kx, (out_rows,) if vec else (out_rows, k), np.float64)
    check(lib.ab_spmm(A.handle(), 1 if trans else 0, px, k, ptr(Y)))
    return Y


def _make_autograd():
    import torch

    class SpMMFn(torch.autograd.Function):
        """sparse_dense_matmul with the TF gradient (oracle: orc_spmm_grad_values / A' dY)."""

        @staticmethod
        def forward(ctx, vv, X, A, trans):
            ctx.A, ctx.trans = A, trans
            ctx.save_for_backward(X)
            return _spmm_raw(A, X.detach(), trans)

        @staticmethod
        def backward(ctx, dY):
            (X,) = ctx.saved_tensors
            A, trans = ctx.A, ctx.trans
            dY = dY.contiguous()
            vec = dY.dim() == 1
            k = 1 if vec else dY.shape[1]
            gv = gx = None
            h = A.handle()
            if ctx.needs_input_grad[0]:
                T = A.query()[3]
                gv = torch.empty(T, dtype=torch.float64, device=dY.device)
                # for A'X the roles of the two dense operands swap: dv[k] = <X[i_k,:], dY[j_k,:]>
                a, b = (X.contiguous(), dY) if trans else (dY, X.contiguous())
                check(lib.ab_spmm_grad_values(h, ptr(a), ptr(b), k, ptr(gv)))
            if ctx.needs_input_grad[1]:
                gx = _spmm_raw(A, dY, not trans)
            return gv, gx, None, None

    class SolveFn(torch.autograd.Function):
        """sparse_solver / sparse_solver_grad (SparseSolver.h:13-70)."""

        @staticmethod
        def forward(ctx, vv, f, A, method):
            u = _solve_raw(A, f.detach(), method)
            ctx.A, ctx.method = A, method
            ctx.save_for_backward(u)
            return u

        @staticmethod
        def backward(ctx, du):
            (u,) = ctx.saved_tensors
            A = ctx.A
            du = du.contiguous()
            T = A.query()[3]
            gv = torch.empty(T, dtype=torch.float64, device=du.device) if ctx.needs_input_grad[0] else None
            gf = torch.empty_like(u)
            check(lib.ab_solve_grad(A.handle(), ctx.method.encode(), ptr(du), ptr(u), ptr(gv), ptr(gf),
                                    options.sparse.tol, options.sparse.maxiter, None, None))
            return gv, (gf if ctx.needs_input_grad[1] else None), None, None

    return SpMMFn, SolveFn


_AUTOGRAD = None


def _autograd():
    global _AUTOGRAD
    if _AUTOGRAD is None:
        _AUTOGRAD = _make_autograd()
    return _AUTOGRAD


def _needs_grad(*xs) -> bool:
    return any(_is_torch(x) and x.requires_grad for x in xs)


def mul(A: SparseTensor, o):
    """s * o: SpMV for rank-1 `o`, SpMM for rank-2 (sparse.jl:227-241)."""
    if _needs_grad(A.vv, o):
        import torch

        if not _is_torch(o):
            o = torch.as_tensor(np.asarray(o, dtype=np.float64), device=A.vv.device)
        vv = A.vv if _is_torch(A.vv) else torch.as_tensor(np.asarray(A.vv), device=o.device)
        if A._h is not None and _is_torch(A.vv):
            A._refresh_values(A.vv)
        return _autograd()[0].apply(vv, o, A, False)
    return _spmm_raw(A, o, False)


def _grad_operands(A: SparseTensor, o):
    """Tensors for the autograd path; values changed in place (an optimiser step on a parameter) are pushed to the
    cached CSR handle first, as mul() does — otherwise x*A would multiply with stale assembled values."""
    import torch

    if not _is_torch(o):
        o = torch.as_tensor(np.asarray(o, dtype=np.float64), device=A.vv.device)
    vv = A.vv if _is_torch(A.vv) else torch.as_tensor(np.asarray(A.vv), device=o.device)
    if A._h is not None and _is_torch(A.vv):
        A._refresh_values(A.vv)
    return vv, o


def rmul(o, A: SparseTensor):
    """o * s = (s' o')' (sparse.jl:247-253)."""
    if len(_shape(o)) == 1:
        if _needs_grad(A.vv, o):
            vv, o = _grad_operands(A, o)
            return _autograd()[0].apply(vv, o, A, True)
        return _spmm_raw(A, o, True)
    if _needs_grad(A.vv, o):
        vv, o = _grad_operands(A, o)
        return _autograd()[0].apply(vv, o.t().contiguous(), A, True).t()
    ot = o.t().contiguous() if _is_torch(o) else np.ascontiguousarray(np.asarray(o).T)
    r = _spmm_raw(A, ot, True)
    return r.t() if _is_torch(r) else r.T


# ---------------------------------------------------------------------- A \ f
def _resolve_method(method: str) -> str:
    return options.sparse.solver if method == "auto" else method


def _solve_raw(A: SparseTensor, f, method: str):
    if A.m != A.n:                       # sparse.jl:418-431: non-square `\` is the least-squares solve (batched over rows of f)
        from . import structure

        return structure.least_square_raw(A, f)
    if len(_shape(f)) != 1:
        raise ValueError("sparse_solver: f must be rank 1 (SparseSolver.cpp:38)")
    if int(_shape(f)[0]) != A.m:
        raise ValueError("dimension mismatch")
    kf, pf = as_array(f, np.float64)
    u = empty_like_kind(kf, (A.n,), np.float64)
    check(lib.ab_solve(A.handle(), method.encode(), pf, ptr(u), options.sparse.tol, options.sparse.maxiter, None, None))
    return u


def backslash(A, f, method: str = "auto"):
    """A \\ f (sparse.jl:413-445).  A may also be a factorized tuple (sparse.jl:699)."""
    if isinstance(A, tuple):
        return solve(A, f)
    method = _resolve_method(method)
    if A.m != A.n and _needs_grad(A.vv, f):
        from . import structure

        if A._h is not None and _is_torch(A.vv):
            A._refresh_values(A.vv)
        return structure.least_square_autograd(A, f)
    if _needs_grad(A.vv, f):
        import torch

        if not _is_torch(f):
            f = torch.as_tensor(np.asarray(f, dtype=np.float64), device=A.vv.device)
        vv = A.vv if _is_torch(A.vv) else torch.as_tensor(np.asarray(A.vv), device=f.device)
        if A._h is not None and _is_torch(A.vv):
            A._refresh_values(A.vv)
        return _autograd()[1].apply(vv, f, A, method)
    return _solve_raw(A, f, method)


# ---------------------------------------------------------------------- SparseAssembler
def SparseAssembler(handle: int, n: int, tol: float = 0.0) -> int:
    """sparse.jl:480-486 -> create_sparse_assembler (Impl.cpp:61-78)."""
    check(lib.ab_asm_create(int(handle), int(n), float(tol)))
    return int(handle)


def accumulate(handle: int, row: int, cols, vals) -> int:
    """sparse.jl:505-512 -> SparseAccum::push_back (Impl.cpp:11-27).  Returns the handle (the `op`)."""
    c = np.ascontiguousarray(cols, dtype=np.int32).ravel()
    v = np.ascontiguousarray(vals, dtype=np.float64).ravel()
    if c.size != v.size:
        raise ValueError("cols and vals must have the same length")
    check(lib.ab_asm_add(int(handle), int(row), ptr(c), ptr(v), int(v.size)))
    return int(handle)


def assemble(m: int, n: int, ops) -> SparseTensor:
    """sparse.jl:527-534 -> copy_to + destroy (Impl.cpp:37-48,107-118) -> SparseTensor(ii,jj,vv,m,n)."""
    hs = np.atleast_1d(np.asarray(ops)).astype(np.int64)
    h = int(hs[0])
    if np.any(hs != h):
        raise ValueError("assemble: ops from different SparseAssembler handles")
    k = lib.ab_asm_nnz(h)
    if k < 0:
        check(int(k))
    r, c, v = np.empty(k, np.int32), np.empty(k, np.int32), np.empty(k, np.float64)
    check(lib.ab_asm_copy(h, ptr(r), ptr(c), ptr(v)))
    return SparseTensor(r.astype(np.int64), c.astype(np.int64), v, m, n)


# ---------------------------------------------------------------------- compress
def compress_indices(indices, v):
    """sparse_compress op (SparseCompress.cpp:14-38): 0-based [N,2] sorted indices -> merged."""
    ki, pi = as_array(indices, np.int64)
    kv, pv = as_array(v, np.float64)
    N = _numel(kv)
    nout = C.c_int64(0)
    check(lib.ab_compress_count(N, pi, C.byref(nout)))
    oi = empty_like_kind(kv, (nout.value, 2), np.int64)
    ov = empty_like_kind(kv, (nout.value,), np.float64)
    check(lib.ab_compress(N, pi, pv, ptr(oi), ptr(ov)))
    return oi, ov


def compress_grad(indices, grad_nv):
    ki, pi = as_array(indices, np.int64)
    kg, pg = as_array(grad_nv, np.float64)
    N = _numel(ki) // 2
    out = empty_like_kind(kg, (N,), np.float64)
    check(lib.ab_compress_grad(N, pi, pg, ptr(out)))
    return out


def compress(A: SparseTensor) -> SparseTensor:
    """compress(A) (sparse.jl:776-781): merge duplicated indices of a SORTED SparseTensor."""
    ii, jj, vv = find(A)
    idx = np.stack([np.asarray(ii) - 1, np.asarray(jj) - 1], axis=1).astype(np.int64)
    oi, ov = compress_indices(idx, vv)
    oi = np.asarray(oi.cpu() if _is_torch(oi) else oi)
    return SparseTensor(oi[:, 0] + 1, oi[:, 1] + 1, ov, A.m, A.n)


# ---------------------------------------------------------------------- factorize / solve
def factorize(A, max_cache_size: int = 999999):
    """factorize (sparse.jl:671-681): returns (A, id); the id keys the native cache."""
    if not isinstance(A, SparseTensor):
        A = SparseTensor(A)
    ii, jj, vv = find(A)
    ki, pi = as_array(ii, np.int64)
    kj, pj = as_array(jj, np.int64)
    kv, pv = as_array(vv, np.float64)
    o = C.c_int64(0)
    check(lib.ab_factorize(pi, pj, pv, _numel(kv), A.m, int(max_cache_size), C.byref(o)))
    return (A, o.value)


def solve(A_factorized, rhs):
    """solve (sparse.jl:688-694) -> Solve.h:3-14."""
    A, o = A_factorized
    kr, pr = as_array(rhs, np.float64)
    out = empty_like_kind(kr, (A.n,), np.float64)
    check(lib.ab_factorized_solve(int(o), pr, A.m, ptr(out)))
    return out
