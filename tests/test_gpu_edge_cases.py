"""Edge cases of the hot path on the GPU: empty and degenerate inputs, all-duplicate streams, a row far
longer than any tile (general-kernel fallbacks inside the solver), boundary-row lists that repeat or cover
everything, non-finite values (never a silent NaN), and handle misuse."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu

E = np.zeros(0, np.int64)
Ev = np.zeros(0)


def test_empty_matrix_everywhere(pkg):
    A = pkg.SparseTensor(E, E, Ev, 5, 7)
    rp, ci, va = A.csr()
    assert rp.tolist() == [0] * 6 and ci.size == 0 and va.size == 0 and A.query() == (5, 7, 0, 0)
    assert np.array_equal(A @ np.ones(7), np.zeros(5)) and np.array_equal(np.ones(5) @ A, np.zeros(7))
    assert np.array_equal(A @ np.ones((7, 3)), np.zeros((5, 3)))
    assert np.array_equal(pkg.to_dense(A), np.zeros((5, 7)))
    i, j, v = pkg.find(A)
    assert len(i) == len(j) == len(v) == 0
    B = pkg.SparseTensor(np.array([2]), np.array([3]), np.array([4.0]), 5, 7)
    assert np.array_equal((A + B).toarray(), B.toarray()) and np.array_equal((B - A).toarray(), B.toarray())
    assert pkg.zero_out_row(A, [1, 2]).query()[2] == 0
    assert pkg.getindex(B, E, E).shape == (0, 0)
    assert pkg.getindex(B, [2], [3]).toarray().tolist() == [[4.0]]
    assert np.array_equal(pkg.vcat(A, B).toarray(), np.vstack([np.zeros((5, 7)), B.toarray()]))
    assert (pkg.SparseTensor(E, E, Ev, 3, 5) @ A).query()[:3] == (3, 7, 0)          # SpGEMM with empty operands
    S = pkg.scatter_update(B, [2], [3], pkg.SparseTensor(E, E, Ev, 1, 1))            # A[2,3] = empty block: entry removed
    assert S.query()[2] == 0
    oi, ov = pkg.compress_indices(np.zeros((0, 2), np.int64), Ev)
    assert oi.shape == (0, 2) and ov.size == 0


def test_all_duplicates_and_single_entry(pkg):
    T = 5000
    vv = np.random.default_rng(233).standard_normal(T)
    A = pkg.SparseTensor(np.full(T, 3), np.full(T, 2), vv, 4, 4)
    rp, ci, va = A.csr()
    s = 0.0
    for x in vv:                       # left-to-right, like setFromTriplets
        s += x
    assert rp.tolist() == [0, 0, 0, 1, 1] and ci.tolist() == [1] and va[0] == s
    assert np.array_equal(A.slot_map(), np.zeros(T, np.int32))
    g = np.empty(T)
    x, dy = np.arange(1.0, 5.0), np.arange(2.0, 6.0)
    pkg.check(pkg.lib.ab_spmm_grad_values(A.handle(), dy.ctypes.data, x.ctypes.data, 1, g.ctypes.data))
    assert np.all(g == dy[2] * x[1])   # every duplicate receives the same gradient
    one = pkg.SparseTensor(np.array([1]), np.array([1]), np.array([4.0]), 1, 1)
    assert pkg.backslash(one, np.array([2.0]), "SparseLU").tolist() == [0.5]
    assert pkg.backslash(one, np.array([0.0]), "CG").tolist() == [0.0]               # b = 0 -> x = 0


def test_long_row_falls_back_to_general_kernels_in_the_solver(pkg, orc):
    """Arrowhead SPD matrix: one row/column with n entries — no tile, no dictionary; CG and the legacy strings
    must go through the vector kernel and still meet the residual bar."""
    n = 30000
    rng = np.random.default_rng(233)
    w = rng.uniform(0.1, 1.0, n - 1) * 1e-2
    d = rng.uniform(1.0, 2.0, n)
    d[0] += float(np.sum(np.abs(w))) + 1.0
    ii = np.concatenate([np.arange(1, n + 1), np.ones(n - 1, np.int64), np.arange(2, n + 1)])
    jj = np.concatenate([np.arange(1, n + 1), np.arange(2, n + 1), np.ones(n - 1, np.int64)])
    vv = np.concatenate([d, w, w])
    A = pkg.SparseTensor(ii, jj, vv, n, n)
    S = sp.coo_matrix((vv, (ii - 1, jj - 1)), shape=(n, n)).tocsr()
    x = rng.standard_normal(n)
    assert np.max(np.abs(A @ x - S @ x)) <= 1e-12 * np.max(np.abs(S @ x))
    f = rng.uniform(size=n)
    info = C.c_double(0)
    for method in ("CG", "SparseLU", "BiCGSTAB"):
        u = pkg.backslash(A, f, method)
        assert np.linalg.norm(S @ u - f) / np.linalg.norm(f) <= 1e-10, method
    pkg.check(pkg.lib.ab_csr_get_info(A.handle(), b"tile_rows", C.byref(info)))
    assert info.value == 0                                                           # not a tile-kernel matrix
    pkg.check(pkg.lib.ab_csr_get_info(A.handle(), b"npat", C.byref(info)))
    assert info.value <= 0


def test_zero_out_row_list_quirks(pkg, orc):
    rng = np.random.default_rng(5)
    ii, jj = rng.integers(1, 21, 400), rng.integers(1, 21, 400)
    vv = rng.standard_normal(400)
    ind = np.stack([ii - 1, jj - 1], 1).astype(np.int64)
    op = pkg.load_system_op("zero_out_row")
    for bd in ([7, 7, 3, 7], [20, 1], list(range(1, 21)), [25, 2]):                   # repeats, unsorted, everything, beyond m
        bd = np.array(bd, np.int64)
        oi, ov = op(ind, vv, bd)
        oi0, ov0 = orc.zero_out_row(ind, vv, bd)
        assert np.array_equal(oi, oi0) and np.array_equal(ov, ov0)
    with pytest.raises(pkg.ADCMEError) as e:
        op(ind, vv, np.array([0], np.int64))                                          # 1-based: 0 is out of range
    assert e.value.code == -4


def test_non_finite_values_are_reported_not_returned(pkg, orc):
    ii, jj, vv = orc.facewise_triplets_2d(12, seed=233)
    vv = vv.copy()
    vv[5] = np.nan
    A = pkg.SparseTensor(ii, jj, vv, 144, 144)
    for method in ("CG", "SparseLU"):
        with pytest.raises(pkg.ADCMEError) as e:
            pkg.backslash(A, np.ones(144), method)
        assert e.value.code in (-5, -6)
    B = pkg.SparseTensor(*orc.facewise_triplets_2d(12, seed=233), 144, 144)
    f = np.ones(144)
    f[3] = np.inf
    with pytest.raises(pkg.ADCMEError):
        pkg.backslash(B, f, "CG")


def test_handle_misuse_and_shape_errors(pkg):
    assert pkg.lib.ab_csr_destroy(987654321) == -2
    assert pkg.lib.ab_trip_size(424242) == -2 and pkg.lib.ab_trip_destroy(424242) == -2
    h = C.c_int64(0)
    assert pkg.lib.ab_spgemm(111, 222, 0, C.byref(h)) == -2
    A = pkg.SparseTensor(np.array([1]), np.array([2]), np.array([1.0]), 2, 3)
    one2, out2 = np.ones(2), np.ones(2)
    with pytest.raises(pkg.ADCMEError):
        pkg.check(pkg.lib.ab_solve(A.handle(), b"CG", one2.ctypes.data, out2.ctypes.data, 0.0, 0, None, None))
    with pytest.raises(pkg.ADCMEError):
        pkg.check(pkg.lib.ab_solve(A.handle(), b"NoSuchSolver", one2.ctypes.data, out2.ctypes.data, 0.0, 0, None, None))
    with pytest.raises(ValueError):
        A @ np.ones(2)
    hc = C.c_int64(0)
    B = pkg.SparseTensor(np.array([1]), np.array([1]), np.array([1.0]), 3, 3)       # keep alive: its finaliser frees the handle
    rc = pkg.lib.ab_csr_axpby(A.handle(), 1.0, B.handle(), 1.0, C.byref(hc))
    assert rc == -1 and b"does not match" in pkg.lib.ab_last_error()
