"""GPU parity tests for the structural ops between assembly and solve and the
remaining SparseTensor algebra (N4): the CUDA path through the C-ABI vs the CPU oracle's line-by-line
restatements of the reference loops, on the same seeded inputs.  Integer outputs (indices, order)
and moved values are bit-exact; SpGEMM sums follow the reference's k-ascending order.
"""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def rand_coo(rng, m, n, T, dup_frac=0.0):
    ii, jj = rng.integers(1, m + 1, T), rng.integers(1, n + 1, T)
    k = int(T * dup_frac)
    if k:
        src, dst = rng.integers(0, T, k), rng.integers(0, T, k)
        ii[dst], jj[dst] = ii[src], jj[src]
    return ii.astype(np.int64), jj.astype(np.int64), rng.standard_normal(T)


def unique_coo(rng, m, n, density):
    A = sp.random(m, n, density, random_state=int(rng.integers(1 << 30)), format="coo")
    p = rng.permutation(A.nnz)
    return (A.row[p] + 1).astype(np.int64), (A.col[p] + 1).astype(np.int64), A.data[p].copy()


def dense(ii, jj, vv, m, n):
    D = np.zeros((m, n))
    np.add.at(D, (np.asarray(ii) - 1, np.asarray(jj) - 1), np.asarray(vv))
    return D


# ======================================================================== zero_out_row
@pytest.mark.parametrize("m,n,T,nbd", [(10, 10, 30, 3), (500, 400, 20000, 57), (3000, 3000, 100000, 1500), (7, 7, 0, 2),
                                       (50, 50, 300, 0)])
def test_zero_out_row_matches_oracle(pkg, orc, m, n, T, nbd):
    rng = np.random.default_rng(233)
    ii, jj, vv = rand_coo(rng, m, n, T, dup_frac=0.2)
    bd = rng.choice(np.arange(1, m + 1), size=nbd, replace=False).astype(np.int64)
    ind = np.stack([ii - 1, jj - 1], axis=1)
    op = pkg.load_system_op("zero_out_row")
    oind, ovv = op(ind, vv, bd)
    oind0, ovv0 = orc.zero_out_row(ind, vv, bd)
    assert np.array_equal(oind, oind0) and np.array_equal(ovv, ovv0)
    g = rng.standard_normal(len(ovv))
    _, gv, _ = op.grad(g, oind, ovv, ind, vv, bd)
    assert np.array_equal(gv, orc.zero_out_row_grad(ind, bd, g))


def test_zero_out_row_reference_test(pkg, golden):
    z = golden["zero_out_row_shape"]                                   # test/sparse.jl:383-389
    A = sp.random(10, 10, 0.3, random_state=7, format="csr")
    B = pkg.zero_out_row(pkg.SparseTensor(A), z["bd"])
    A = A.toarray()
    A[np.array(z["bd"]) - 1, :] = 0
    assert np.linalg.norm(B.toarray() - A) < 1e-8


def test_zero_out_row_device_tensors_and_autograd(pkg, orc):
    import torch

    rng = np.random.default_rng(5)
    ii, jj, vv = rand_coo(rng, 200, 200, 5000)
    bd = np.array([1, 17, 200, 64], dtype=np.int64)
    vt = torch.tensor(vv, device="cuda", requires_grad=True)
    A = pkg.SparseTensor(torch.tensor(ii, device="cuda"), torch.tensor(jj, device="cuda"), vt, 200, 200)
    B = pkg.zero_out_row(A, bd)
    assert B.vv.is_cuda
    w = torch.tensor(rng.standard_normal(B.vv.numel()), device="cuda")
    (B.vv * w).sum().backward()
    g0 = orc.zero_out_row_grad(np.stack([ii - 1, jj - 1], 1), bd, w.cpu().numpy())
    assert np.array_equal(vt.grad.cpu().numpy(), g0)


# ======================================================================== scatter_update / scatter_add
@pytest.mark.parametrize("m,n,ni,nj,dens", [(10, 10, 3, 3, 0.3), (300, 200, 40, 25, 0.05), (2000, 2000, 500, 300, 0.002)])
def test_scatter_update_matches_oracle(pkg, orc, m, n, ni, nj, dens):
    rng = np.random.default_rng(233)
    oii, ojj, ovv = unique_coo(rng, m, n, dens)
    uii, ujj, uvv = unique_coo(rng, ni, nj, 0.5)
    ii = rng.choice(np.arange(1, m + 1), ni, replace=False).astype(np.int64)     # unsorted, unique
    jj = rng.choice(np.arange(1, n + 1), nj, replace=False).astype(np.int64)
    op = pkg.load_system_op("sparse_scatter_update")
    I, J, V = op(oii, ojj, ovv, m, n, uii, ujj, uvv, ii, jj)
    I0, J0, V0 = orc.sparse_scatter_update(oii, ojj, ovv, uii, ujj, uvv, m, n, ii, jj)
    assert np.array_equal(I, I0) and np.array_equal(J, J0) and np.array_equal(V, V0)
    g = rng.standard_normal(len(V))
    r = op.grad(g, I, J, V, oii, ojj, ovv, m, n, uii, ujj, uvv, ii, jj)
    g1, g2 = orc.sparse_scatter_update_grad(g, oii, ojj, len(uvv), ii, jj)
    assert np.array_equal(r[2], g1) and np.array_equal(r[7], g2)
    assert all(r[k] is None for k in (0, 1, 3, 4, 5, 6, 8, 9))


def test_scatter_update_add_reference_test(pkg, golden):
    g = golden["scatter_update_shape"]                                 # test/sparse.jl:294-308
    A = sp.random(10, 10, 0.3, random_state=11, format="csr")
    B = sp.random(3, 3, 0.6, random_state=12, format="csr")
    ii, jj = np.array(g["ii"]), np.array(g["jj"])
    u = pkg.scatter_update(A, ii, jj, B)
    Cm = A.toarray().copy()
    Cm[np.ix_(ii - 1, jj - 1)] = B.toarray()
    assert np.array_equal(u.toarray(), Cm)
    u = pkg.scatter_add(A, ii, jj, B)
    Cm = A.toarray().copy()
    Cm[np.ix_(ii - 1, jj - 1)] += B.toarray()
    assert np.allclose(u.toarray(), Cm, rtol=1e-15, atol=0)


def test_scatter_update_rejects_out_of_range(pkg):
    op = pkg.load_system_op("sparse_scatter_update")
    one = np.array([1], dtype=np.int64)
    with pytest.raises(pkg.ADCMEError) as e:
        op(one, one, np.array([1.0]), 2, 2, one, one, np.array([1.0]), np.array([3], dtype=np.int64), one)
    assert e.value.code == -4


# ======================================================================== getindex
@pytest.mark.parametrize("m,n,T,ni,nj", [(10, 10, 30, 2, 2), (20, 30, 180, 3, 3), (400, 600, 30000, 120, 77),
                                         (3000, 2500, 200000, 900, 1100)])
def test_sparse_indexing_matches_oracle(pkg, orc, m, n, T, ni, nj):
    rng = np.random.default_rng(233)
    ii1, jj1, vv1 = rand_coo(rng, m, n, T, dup_frac=0.2)               # duplicates: the op sums them
    ii = rng.choice(np.arange(1, m + 1), ni, replace=False).astype(np.int64)
    jj = rng.choice(np.arange(1, n + 1), nj, replace=False).astype(np.int64)
    op = pkg.load_system_op("sparse_indexing")
    I, J, V = op(ii1, jj1, vv1, m, n, ii, jj)
    I0, J0, V0 = orc.sparse_indexing(ii1, jj1, vv1, m, n, ii, jj)
    assert np.array_equal(I, I0) and np.array_equal(J, J0)             # same order, bit-exact integers
    assert np.array_equal(V, V0)                                       # duplicate sums in input order
    g = rng.standard_normal(len(V))
    r = op.grad(g, I, J, V, ii1, jj1, vv1, m, n, ii, jj)
    assert np.array_equal(r[2], orc.sparse_indexing_grad(g, I, J, ii1, jj1, ii, jj))


def test_sparse_indexing_grad_with_repeated_indices_is_ordered(pkg, orc):
    """Repeated entries in ii / jj make several output positions point at the same input triplet: their gradient
    contributions are summed in ascending output order (the reference's sequential loop, SparseIndexing.h:66-83) by a
    sort + fixed-order segment reduction — no floating-point atomics — so the result equals the oracle bit for bit and is
    the same on every run."""
    rng = np.random.default_rng(7)
    m, n, T = 300, 400, 20000
    ii1, jj1, vv1 = rand_coo(rng, m, n, T, dup_frac=0.2)
    ii = rng.choice(np.arange(1, m + 1), 150, replace=True).astype(np.int64)      # repeats
    jj = rng.choice(np.arange(1, n + 1), 120, replace=True).astype(np.int64)
    I0, J0, V0 = orc.sparse_indexing(ii1, jj1, vv1, m, n, ii, jj)
    g = rng.standard_normal(len(V0)) * 10.0 ** rng.integers(-8, 8, len(V0))       # wide range: the order of the sum matters
    op = pkg.load_system_op("sparse_indexing")
    r1 = op.grad(g, I0, J0, V0, ii1, jj1, vv1, m, n, ii, jj)[2]
    r2 = op.grad(g, I0, J0, V0, ii1, jj1, vv1, m, n, ii, jj)[2]
    assert np.array_equal(r1, r2)
    assert np.array_equal(r1, orc.sparse_indexing_grad(g, I0, J0, ii1, jj1, ii, jj))


def test_getindex_reference_tests(pkg, golden):
    B1 = sp.random(10, 10, 0.3, random_state=13, format="csr")         # test/sparse.jl:63-69
    B = pkg.SparseTensor(B1)
    assert np.array_equal(B[2:3, 2:3].toarray(), B1[1:3, 1:3].toarray())
    assert np.array_equal(B[2:3, :].toarray(), B1[1:3, :].toarray())
    assert np.array_equal(B[:, 2:3].toarray(), B1[:, 1:3].toarray())
    g = golden["get_index_bool_mask"]                                  # test/sparse.jl:316-321
    M = pkg.spdiag(np.array(g["diag"]))
    idof = np.array(g["idof"])
    assert M[idof, idof].toarray().tolist() == g["expect_dense"]
    A = sp.random(20, 30, 0.3, random_state=14, format="csr")          # test/sparse.jl:181-190
    i1, j1 = np.array([7, 2, 19]), np.array([30, 4, 11])
    assert np.array_equal(pkg.SparseTensor(A)[i1, j1].toarray(), A[i1 - 1][:, j1 - 1].toarray())
    assert np.array_equal(pkg.SparseTensor(A)[3, :], A[2, :].toarray().ravel())   # integer index squeezes


def test_getindex_autograd(pkg, orc):
    import torch

    rng = np.random.default_rng(3)
    ii1, jj1, vv1 = rand_coo(rng, 60, 50, 900, dup_frac=0.3)
    vt = torch.tensor(vv1, device="cuda", requires_grad=True)
    A = pkg.SparseTensor(ii1, jj1, vt, 60, 50)
    ii, jj = np.array([5, 1, 33, 60]), np.array([50, 2, 9])
    S = A[ii, jj]
    w = torch.tensor(rng.standard_normal(S.vv.numel()), device="cuda")
    (S.vv * w).sum().backward()
    g0 = orc.sparse_indexing_grad(w.cpu().numpy(), np.asarray(S.ii.cpu()), np.asarray(S.jj.cpu()), ii1, jj1, ii, jj)
    assert np.array_equal(vt.grad.cpu().numpy(), g0)


# ======================================================================== A + B, spdiag, spzero
def test_add_sub_reference_tests(pkg):
    rng = np.random.default_rng(233)
    A1 = sp.random(10, 5, 0.3, random_state=15, format="csr")
    B1 = sp.random(10, 5, 0.3, random_state=16, format="csr")
    A, B = pkg.SparseTensor(A1), pkg.SparseTensor(B1)
    assert np.array_equal((A + B).toarray(), (A1 + B1).toarray())
    assert np.array_equal((A - B).toarray(), (A1 - B1).toarray())
    assert np.array_equal((-B).toarray(), -B1.toarray())               # test/sparse.jl:23-34
    D = rng.uniform(size=(10, 5))
    assert np.array_equal(D + B, D + B1.toarray()) and np.array_equal(B - D, B1.toarray() - D)
    with pytest.raises(ValueError):
        A + pkg.SparseTensor(sp.random(5, 5, 0.3, random_state=1, format="csr"))


def test_axpby_on_device_handles(pkg, orc):
    rng = np.random.default_rng(233)
    ii, jj, vv = rand_coo(rng, 300, 200, 6000, dup_frac=0.2)
    i2, j2, v2 = rand_coo(rng, 300, 200, 4000, dup_frac=0.2)
    A, B = pkg.SparseTensor(ii, jj, vv, 300, 200), pkg.SparseTensor(i2, j2, v2, 300, 200)
    h = C.c_int64(0)
    pkg.check(pkg.lib.ab_csr_axpby(A.handle(), 2.0, B.handle(), -0.5, C.byref(h)))
    S = pkg.SparseTensor.from_handle(h.value)
    ra, ca, va = A.csr()
    rb, cb, vb = B.csr()
    A0 = sp.csr_matrix((va, ca, ra), shape=(300, 200))
    B0 = sp.csr_matrix((vb, cb, rb), shape=(300, 200))
    ref = (2.0 * A0 + (-0.5) * B0).toarray()
    assert np.array_equal(S.toarray(), ref)
    # union pattern, structural zeros kept
    assert S.query()[2] == (abs(A0) + abs(B0)).nnz


def test_spdiag_spzero_reference_tests(pkg, golden):
    rng = np.random.default_rng(233)
    p = rng.uniform(size=10)
    Bd = pkg.spdiag(p)                                                 # test/sparse.jl:164-172
    assert Bd._diag and np.array_equal(Bd.toarray(), np.diag(p))
    assert np.array_equal(pkg.spdiag(10).toarray(), np.eye(10))
    assert np.array_equal(pkg.spzero(10).toarray(), np.zeros((10, 10)))   # test/sparse.jl:174-179
    assert np.array_equal(pkg.spzero(10, 20).toarray(), np.zeros((10, 20)))
    g = golden["spdiag_pairs"]                                         # test/sparse.jl:207-258
    a, b = rng.uniform(size=10), rng.uniform(size=7)
    A = pkg.spdiag(g["n"], (0, a), (-3, b), (3, 4 * b))
    assert np.array_equal(A.toarray(), np.diag(a) + np.diag(b, -3) + np.diag(4 * b, 3))


# ======================================================================== SpGEMM
@pytest.mark.parametrize("m,n,k,da,db", [(10, 5, 20, 0.3, 0.3), (200, 300, 150, 0.05, 0.04), (3000, 2000, 2500, 0.004, 0.003),
                                         (4, 6, 3, 0.0, 0.5)])
def test_spgemm_op_matches_oracle(pkg, orc, m, n, k, da, db):
    rng = np.random.default_rng(233)
    A = sp.random(m, n, da, random_state=21, format="coo")
    B = sp.random(n, k, db, random_state=22, format="coo")
    pa, pb = rng.permutation(A.nnz), rng.permutation(B.nnz)
    a = (A.row[pa].astype(np.int64), A.col[pa].astype(np.int64), A.data[pa])
    b = (B.row[pb].astype(np.int64), B.col[pb].astype(np.int64), B.data[pb])
    op = pkg.load_system_op("sparse_sparse_mat_mul")
    I, J, V = op(*a, *b, m, n, k)
    I0, J0, V0 = orc.sparse_sparse_mat_mul(*a, *b, m, n, k)
    assert np.array_equal(I, I0) and np.array_equal(J, J0)             # column-major emission, structural zeros kept
    assert np.array_equal(V, V0)                                       # k-ascending sums, no FMA contraction on either side
    assert np.allclose(dense(I, J, V, m, k), (A @ B).toarray(), rtol=1e-13, atol=1e-300)


def test_spgemm_duplicates_and_stencil(pkg, orc):
    # A from a face-wise stream (duplicates summed first, as setFromTriplets does), C = A*A
    ii, jj, vv = orc.facewise_triplets_2d(30, seed=233)
    A = pkg.SparseTensor(ii, jj, vv, 900, 900)
    Cs = A @ A
    rp, ci, va = A.csr()
    A0 = sp.csr_matrix((va, ci, rp), shape=(900, 900))
    assert np.array_equal(Cs.toarray(), (A0 @ A0).toarray())           # small integers: exact in any order
    assert Cs.query()[2] == (A0 @ A0).nnz


def test_diag_fast_paths_reference_tests(pkg, orc):
    d = np.array([1.0, 2, 3, 4, 5])                                    # test/sparse.jl:149-161
    B = sp.random(5, 20, 0.3, random_state=23, format="coo")
    A = sp.random(10, 5, 0.5, random_state=24, format="coo")
    Ds = pkg.spdiag(d)
    assert np.array_equal((Ds @ pkg.SparseTensor(B)).toarray(), (sp.diags(d) @ B).toarray())
    assert np.array_equal((pkg.SparseTensor(A) @ Ds).toarray(), (A @ sp.diags(d)).toarray())
    I, J, V = pkg.load_system_op("diag_sparse_mat_mul")(None, None, d, B.row.astype(np.int64), B.col.astype(np.int64),
                                                        B.data, 5, 5, 20)
    I0, J0, V0 = orc.diag_sparse_mat_mul(d, B.row, B.col, B.data, 5, 20)
    assert np.array_equal(I, I0) and np.array_equal(J, J0) and np.array_equal(V, V0)
    I, J, V = pkg.load_system_op("sparse_diag_mat_mul")(A.row.astype(np.int64), A.col.astype(np.int64), A.data, None, None,
                                                        d, 10, 5, 5)
    I0, J0, V0 = orc.sparse_diag_mat_mul(A.row, A.col, A.data, d, 10, 5)
    assert np.array_equal(I, I0) and np.array_equal(J, J0) and np.array_equal(V, V0)


def test_spgemm_size_mismatch_raises(pkg):
    A = pkg.SparseTensor(sp.random(4, 5, 0.5, random_state=1, format="csr"))
    B = pkg.SparseTensor(sp.random(4, 5, 0.5, random_state=2, format="csr"))
    with pytest.raises(ValueError):
        A @ B


# ======================================================================== the FEM idiom end to end
def test_dirichlet_pipeline_assemble_zero_rows_add_identity_solve(pkg, orc):
    """examples/while_loop/fem.jl:68-75 idiom: assemble K, zero the boundary rows, put 1 on their
    diagonal, solve — all device ops; checked against scipy's direct solve."""
    n = 24
    N = n * n
    ii, jj, vv = orc.facewise_triplets_2d(n, seed=233)
    K = pkg.SparseTensor(ii, jj, vv, N, N)
    idx = np.arange(N).reshape(n, n)
    bd = np.unique(np.concatenate([idx[0], idx[-1], idx[:, 0], idx[:, -1]])) + 1
    Kz = pkg.zero_out_row(K, bd)
    Kbc = Kz + pkg.SparseTensor(bd, bd, np.ones(len(bd)), N, N)
    rng = np.random.default_rng(233)
    f = rng.uniform(size=N)
    u = pkg.backslash(Kbc, f, "SparseLU")
    rp, ci, va = K.csr()
    K0 = sp.csr_matrix((va, ci, rp), shape=(N, N)).tolil()
    K0[bd - 1, :] = 0
    K0[bd - 1, bd - 1] = 1.0
    import scipy.sparse.linalg as spla

    u0 = spla.spsolve(K0.tocsc(), f)
    assert np.linalg.norm(u - u0) <= 1e-9 * np.linalg.norm(u0)
    assert np.allclose(u[bd - 1], f[bd - 1], rtol=1e-10)


# ======================================================================== vcat / hcat / hvcat, Array, sum
def test_concat_reference_tests(pkg, orc):
    B1 = sp.random(10, 3, 0.3, random_state=31, format="csr")          # test/sparse.jl:56-61
    B = pkg.SparseTensor(B1)
    assert np.array_equal(pkg.vcat(B, B).toarray(), sp.vstack([B1, B1]).toarray())
    assert np.array_equal(pkg.hcat(B, B).toarray(), sp.hstack([B1, B1]).toarray())
    A = sp.random(10, 5, 0.3, random_state=32, format="csr")           # test/sparse.jl:260-267
    Bm = sp.random(10, 5, 0.2, random_state=33, format="csr")
    Cm = sp.random(5, 10, 0.4, random_state=34, format="csr")
    D = pkg.hvcat((2, 1), pkg.SparseTensor(A), pkg.SparseTensor(Bm), pkg.SparseTensor(Cm))
    assert np.array_equal(D.toarray(), sp.vstack([sp.hstack([A, Bm]), Cm]).toarray())
    # op level: exact order and integers
    rng = np.random.default_rng(233)
    i1, j1, v1 = rand_coo(rng, 40, 30, 500, 0.2)
    i2, j2, v2 = rand_coo(rng, 25, 30, 300, 0.2)
    op = pkg.load_system_op("sparse_concate")
    for hc, (a, b) in ((False, ((i1, j1, v1, 40, 30), (i2, j2, v2, 25, 30))),):
        I, J, V = op(*a, *b, hc)
        I0, J0, V0 = orc.sparse_concate(a[0], a[1], a[2], a[3], a[4], b[0], b[1], b[2], hc)
        assert np.array_equal(I, I0) and np.array_equal(J, J0) and np.array_equal(V, V0)
    with pytest.raises(ValueError):
        pkg.hcat(pkg.SparseTensor(A), pkg.SparseTensor(Cm))


def test_to_dense_and_sum(pkg, orc):
    rng = np.random.default_rng(233)
    ii, jj, vv = rand_coo(rng, 37, 53, 900, dup_frac=0.3)
    A = pkg.SparseTensor(ii, jj, vv, 37, 53)
    D = pkg.to_dense(A)
    D0 = orc.sparse_to_dense(np.stack([ii - 1, jj - 1], 1), vv, 37, 53)
    assert np.array_equal(D, D0)                                       # duplicates summed in input order
    g = rng.standard_normal((37, 53))
    gv = np.empty(len(vv))
    ij = np.ascontiguousarray(np.stack([ii - 1, jj - 1], 1))
    pkg.check(pkg.lib.ab_dense_grad_values(len(vv), ij.ctypes.data, 37, 53, g.ctypes.data, gv.ctypes.data))
    assert np.array_equal(gv, g[ii - 1, jj - 1])                       # SparseToDense.h:8-18
    s = sp.random(10, 20, 0.2, random_state=35, format="csr")          # test/sparse.jl:192-198
    S = pkg.SparseTensor(s)
    assert abs(pkg.spsum(S) - s.sum()) <= 1e-12 * abs(s).sum()
    assert np.allclose(pkg.spsum(S, dims=1), np.asarray(s.sum(axis=0)).ravel(), rtol=1e-13, atol=1e-15)
    assert np.allclose(pkg.spsum(S, dims=2), np.asarray(s.sum(axis=1)).ravel(), rtol=1e-13, atol=1e-15)


# ======================================================================== find / tf.sparse.reorder
def test_find_is_stable_row_major_reorder(pkg):
    import torch

    rng = np.random.default_rng(233)
    A = sp.random(10, 10, 0.3, random_state=41, format="coo")          # test/sparse.jl:269-292
    p = rng.permutation(A.nnz)
    S = pkg.SparseTensor(A.row[p] + 1, A.col[p] + 1, A.data[p], 10, 10)
    i, j, v = pkg.find(S)
    R = A.tocsr().tocoo()                                              # row-major
    assert np.array_equal(i, R.row + 1) and np.array_equal(j, R.col + 1) and np.array_equal(v, R.data)
    assert np.array_equal(pkg.rows(S), i) and np.array_equal(pkg.cols(S), j) and np.array_equal(pkg.values(S), v)
    # duplicates are kept, in input order (stable) — tf.sparse.reorder does not merge
    ii, jj, vv = rand_coo(rng, 50, 40, 3000, dup_frac=0.4)
    i, j, v = pkg.find(pkg.SparseTensor(ii, jj, vv, 50, 40))
    q = np.lexsort((np.arange(len(vv)), jj, ii))
    assert np.array_equal(i, ii[q]) and np.array_equal(j, jj[q]) and np.array_equal(v, vv[q])
    # device-resident values with gradient: the permutation stays on the autograd tape
    vt = torch.tensor(vv, device="cuda", requires_grad=True)
    _, _, v2 = pkg.find(pkg.SparseTensor(ii, jj, vt, 50, 40))
    w = torch.arange(len(vv), dtype=torch.float64, device="cuda")
    (v2 * w).sum().backward()
    g = np.empty(len(vv)); g[q] = np.arange(len(vv), dtype=float)
    assert np.array_equal(vt.grad.cpu().numpy(), g)
    pkg.options.sparse.auto_reorder = False
    try:
        i3, _, _ = pkg.find(pkg.SparseTensor(ii, jj, vv, 50, 40))
        assert np.array_equal(i3, ii)
    finally:
        pkg.reset_default_options()


def test_dense_to_sparse_reference_test(pkg):
    A = sp.random(10, 20, 0.3, random_state=43, format="csr")          # test/sparse.jl:200-205
    B = A.toarray()
    S = pkg.dense_to_sparse(B)
    assert np.array_equal(S.toarray(), B) and S.query()[2] == A.nnz
    R = A.tocoo()
    assert np.array_equal(S.ii, R.row + 1) and np.array_equal(S.jj, R.col + 1) and np.array_equal(S.vv, R.data)   # row-major
    assert np.array_equal(pkg.SparseTensor(B).toarray(), B)            # SparseTensor(A::Array)
    assert pkg.dense_to_sparse(np.zeros((3, 4))).query()[2] == 0
    assert np.array_equal(pkg.to_dense(S), B)


# ======================================================================== trisolve
@pytest.mark.parametrize("n", [1, 2, 3, 10, 1000, 4097, 300000])
def test_trisolve_matches_oracle(pkg, orc, n):
    rng = np.random.default_rng(233)
    a, c = rng.uniform(size=max(n - 1, 0)), rng.uniform(size=max(n - 1, 0))
    b, d = rng.uniform(size=n) + 10, rng.uniform(size=n)
    op = pkg.load_system_op("tri_solve")
    x = op(a, b, c, d)
    x0 = orc.tri_solve(a, b, c, d)
    assert np.max(np.abs(x - x0)) <= 1e-10 * np.max(np.abs(x0))       # cyclic reduction vs Thomas: rounding only
    g = rng.standard_normal(n)
    ga, gb, gc, gd = op.grad(g, x, a, b, c, d)
    ga0, gb0, gc0, gd0 = orc.tri_solve_grad(g, x0, a, b, c)
    for u, v in ((ga, ga0), (gb, gb0), (gc, gc0), (gd, gd0)):
        assert u.shape == v.shape and (v.size == 0 or np.max(np.abs(u - v)) <= 1e-10 * np.max(np.abs(v)))


def test_trisolve_reference_test_autograd_and_errors(pkg):
    import torch

    rng = np.random.default_rng(233)
    n = 10                                                             # test/sparse.jl:352-364
    a, b, c, d = rng.uniform(size=n - 1), rng.uniform(size=n) + 10, rng.uniform(size=n - 1), rng.uniform(size=n)
    A = np.diag(b) + np.diag(a, -1) + np.diag(c, 1)
    u = pkg.trisolve(a, b, c, d)
    assert np.linalg.norm(u - np.linalg.solve(A, d)) < 1e-12
    bt = torch.tensor(b, device="cuda", requires_grad=True)
    dt = torch.tensor(d, device="cuda", requires_grad=True)
    x = pkg.trisolve(a, bt, c, dt)
    x.sum().backward()
    lam = np.linalg.solve(A.T, np.ones(n))
    assert np.allclose(dt.grad.cpu().numpy(), lam, rtol=1e-10) and np.allclose(bt.grad.cpu().numpy(), -lam * u, rtol=1e-10)
    with pytest.raises(ValueError):
        pkg.trisolve(a, b, c[:-1], d)
    with pytest.raises(pkg.ADCMEError) as e:
        pkg.trisolve(np.zeros(2), np.zeros(3), np.zeros(2), np.ones(3))       # singular: zero pivot
    assert e.value.code == -6


# ======================================================================== least squares  A \ f,  A tall
def test_least_square_golden_parity_and_autograd(pkg, orc, golden):
    import torch

    g = golden["least_squares_adjacent"]                               # test/sparse.jl:131-139
    A = pkg.SparseTensor(np.array(g["ii"]), np.array(g["jj"]), np.array(g["vv"], float), g["m"], g["n"])
    o = pkg.backslash(A, np.array(g["ff"], float))
    assert np.linalg.norm(o - np.array(g["expect"])) < g["tol"]
    rng = np.random.default_rng(233)
    m, n, T = 300, 40, 2500
    ii, jj, vv = rng.integers(1, m + 1, T), rng.integers(1, n + 1, T), rng.standard_normal(T)
    ii[:n], jj[:n] = np.arange(1, n + 1), np.arange(1, n + 1)
    vv[:n] += 8.0
    F = rng.standard_normal((3, m))
    op = pkg.load_system_op("sparse_least_square")
    U = op(ii, jj, vv, F, n)                                           # batched: one solve per row of F
    U0 = orc.sparse_least_square(ii, jj, vv, F, m, n)
    assert U.shape == (3, n) and np.max(np.abs(U - U0)) <= 1e-9 * np.max(np.abs(U0))
    Gu = rng.standard_normal((3, n))
    r = op.grad(Gu, U0, ii, jj, vv, F, n)
    gvv0, gf0 = orc.sparse_least_square_grad(Gu, U0, ii, jj, vv, F, m, n)
    assert np.max(np.abs(r[2] - gvv0)) <= 1e-8 * np.max(np.abs(gvv0)) and np.max(np.abs(r[3] - gf0)) <= 1e-8 * np.max(np.abs(gf0))
    vt = torch.tensor(vv, device="cuda", requires_grad=True)
    ft = torch.tensor(F[0], device="cuda", requires_grad=True)
    u = pkg.backslash(pkg.SparseTensor(ii, jj, vt, m, n), ft)
    w = torch.tensor(Gu[0], device="cuda")
    (u * w).sum().backward()
    g1, gf1 = orc.sparse_least_square_grad(Gu[0], U0[0], ii, jj, vv, F[0], m, n)
    assert np.max(np.abs(vt.grad.cpu().numpy() - g1)) <= 1e-8 * np.max(np.abs(g1))
    assert np.max(np.abs(ft.grad.cpu().numpy() - gf1)) <= 1e-8 * np.max(np.abs(gf1))
