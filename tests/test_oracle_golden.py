"""Pins the CPU oracle (oracle/adcme_oracle.c + oracle.py) against every golden vector the
reference's own tests hold for the sparse hot path, and against independent
scipy results on seeded random inputs.  CPU only."""
import numpy as np
import scipy.sparse as sp


def _dense_from_triplets(r, c, v, m, n):
    D = np.zeros((m, n))
    np.add.at(D, (np.asarray(r) - 1, np.asarray(c) - 1), v)
    return D


def test_compress_golden(orc, golden):
    g = golden["compress"]                                   # test/sparse.jl:366-381
    oi, ov = orc.compress(np.array(g["indices_1based"]) - 1, g["v"])
    assert oi.tolist() == g["expect_indices_0based"]
    assert ov.tolist() == g["expect_values"]
    gv = orc.compress_grad(np.array(g["indices_1based"]) - 1, [10.0, 20.0, 30.0])
    assert gv.tolist() == [10.0, 10.0, 20.0, 30.0]            # SparseCompress.h:45-54


def test_assembler_golden(orc, golden):
    for key in ("assembler_tol_filter", "assembler_duplicates", "assembler_docstring_example2"):
        g = golden[key]
        acc = orc.Accumulator(g["n"], g["tol"])
        for a in g["adds"]:
            assert acc.add(a["row"], a["cols"], a["vals"]) == 0
        r, c, v = acc.copy()
        assert np.all(np.diff(r) >= 0), "dump must be row by row"
        assert np.allclose(_dense_from_triplets(r, c, v, g["m"], g["ncol"]), np.array(g["expect_dense"], float)), key


def test_assembler_keeps_insertion_order_and_strict_tol(orc):
    acc = orc.Accumulator(3, 0.5)
    acc.add(2, [3, 1, 2], [1.0, 0.5, -0.6])    # 0.5 is NOT > tol -> dropped (strict, Impl.cpp:16)
    acc.add(1, [2], [7.0])
    acc.add(2, [1], [9.0])
    r, c, v = acc.copy()
    assert r.tolist() == [1, 2, 2, 2] and c.tolist() == [2, 3, 2, 1] and v.tolist() == [7.0, 1.0, -0.6, 9.0]
    acc = orc.Accumulator(3, 0.0)
    assert acc.add(4, [1], [1.0]) != 0          # row > n (Impl.cpp:17-22)
    assert acc.nnz() == 0


def test_mpi_create_matrix_golden(orc, golden):
    g = golden["mpi_create_matrix"]                           # test_mpi_create_matrix.jl:5-21
    rows, ncols, cols = orc.coo_to_rows(g["indices"], g["ilower"])
    assert rows.tolist() == g["expect_rows"]
    assert ncols.tolist() == g["expect_ncols"]
    assert cols.tolist() == g["expect_cols"]


def test_constructor_golden_vs_scipy(orc, golden):
    g = golden["sparse_constructor"]                          # test/sparse.jl:5-10
    V = np.random.default_rng(233).uniform(size=6)
    rp, ci, va, slot = orc.coo_to_csr(g["I"], g["J"], V, g["m"], g["n"], base=1)
    A = sp.coo_matrix((V, (np.array(g["I"]) - 1, np.array(g["J"]) - 1)), shape=(g["m"], g["n"])).tocsr()
    A.sort_indices()
    assert np.array_equal(rp, A.indptr) and np.array_equal(ci, A.indices) and np.allclose(va, A.data, rtol=0, atol=0)
    assert np.array_equal(va[slot], V)


def test_coo_to_csr_duplicates_sum_in_input_order(orc):
    # three duplicates whose fp64 sum depends on the order: setFromTriplets adds in input order
    ii, jj = [1, 1, 1, 2], [1, 1, 1, 2]
    vv = [1e16, 1.0, -1e16, 5.0]
    rp, ci, va, slot = orc.coo_to_csr(ii, jj, vv, 2, 2)
    assert va[0] == (1e16 + 1.0) - 1e16 and va[1] == 5.0
    assert slot.tolist() == [0, 0, 0, 1]
    rng = np.random.default_rng(233)
    m, n, T = 37, 29, 2000
    ii, jj, vv = rng.integers(1, m + 1, T), rng.integers(1, n + 1, T), rng.standard_normal(T)
    rp, ci, va, slot = orc.coo_to_csr(ii, jj, vv, m, n)
    A = sp.coo_matrix((vv, (ii - 1, jj - 1)), shape=(m, n)).tocsr()
    A.sum_duplicates(); A.sort_indices()
    assert np.array_equal(rp, A.indptr) and np.array_equal(ci, A.indices)
    assert np.allclose(va, A.data, rtol=1e-13, atol=1e-13)
    # slot maps every triplet to the entry holding its (row, col)
    rows_of_entry = np.repeat(np.arange(m), np.diff(rp))
    assert np.array_equal(rows_of_entry[slot], ii - 1) and np.array_equal(ci[slot], jj - 1)
    try:
        orc.coo_to_csr([3], [1], [1.0], 2, 2)
        assert False
    except IndexError:
        pass


def test_spmm_and_gradients_vs_scipy(orc):
    rng = np.random.default_rng(233)
    m, n, k, T = 23, 17, 5, 200
    idx = np.stack([rng.integers(0, m, T), rng.integers(0, n, T)], 1)
    val = rng.standard_normal(T)
    B = rng.standard_normal((n, k))
    A = sp.coo_matrix((val, (idx[:, 0], idx[:, 1])), shape=(m, n)).tocsr()
    assert np.allclose(orc.spmm_coo(idx, val, m, B), A @ B, rtol=1e-13, atol=1e-13)          # test/sparse.jl:42-54
    assert np.allclose(orc.spmm_coo(idx, val, m, B[:, 0]), A @ B[:, 0], rtol=1e-13, atol=1e-13)
    G = rng.standard_normal((m, k))
    assert np.allclose(orc.spmm_coo(idx, val, n, G, adjoint_a=True), A.T @ G, rtol=1e-13, atol=1e-13)
    gv = orc.spmm_grad_values(idx, G, B)
    assert np.allclose(gv, np.einsum("tk,tk->t", G[idx[:, 0]], B[idx[:, 1]]), rtol=1e-13, atol=1e-13)
    # finite-difference check of the value gradient of L = sum(G * (A B))
    e = np.zeros(T); e[7] = 1e-6
    L = lambda v: np.sum(G * orc.spmm_coo(idx, v, m, B))
    assert abs((L(val + e) - L(val - e)) / 2e-6 - gv[7]) < 1e-7


def test_solver_forward_backward_known_answers(orc, golden):
    rng = np.random.default_rng(233)
    # A\b vs dense solve (test/sparse.jl:71-78: I + sprand(10,10,0.1))
    d = 10
    A = sp.identity(d) + sp.random(d, d, 0.1, random_state=233)
    coo = A.tocoo()
    ii, jj, vv = coo.row + 1, coo.col + 1, coo.data
    b = rng.uniform(size=d)
    u = orc.sparse_solver(ii, jj, vv, b)
    assert np.allclose(u, np.linalg.solve(A.toarray(), b), rtol=1e-12, atol=1e-12)
    # both code paths of the oracle (dense LU / SuperLU) agree
    u2 = orc.sparse_solver(ii, jj, vv, b, dense_limit=0)
    assert np.allclose(u, u2, rtol=1e-12, atol=1e-12)
    # gradient formulas (SparseSolver.h:46-70) vs finite differences of L = <g, A^{-1} b>
    g = rng.standard_normal(d)
    gvv, gf = orc.sparse_solver_grad(g, u, ii, jj, vv)
    L = lambda v, f: g @ orc.sparse_solver(ii, jj, v, f)
    for k in (0, len(vv) // 2, len(vv) - 1):
        e = np.zeros(len(vv)); e[k] = 1e-6
        assert abs((L(vv + e, b) - L(vv - e, b)) / 2e-6 - gvv[k]) < 1e-6
    e = np.zeros(d); e[3] = 1e-6
    assert abs((L(vv, b + e) - L(vv, b - e)) / 2e-6 - gf[3]) < 1e-8
    # singular matrix must fail (test/sparse.jl:335-339)
    for lim in (600, 0):
        try:
            orc.sparse_solver(np.array([1]), np.array([1]), np.array([0.0]), rng.uniform(size=d), d=d, dense_limit=lim)
            assert False, "singular solve must raise"
        except orc.SolverError:
            pass
    # Newton + adjoint known answer (test/optim.jl:235-249): x^3 = theta, J = 3 diag(x^2)
    n = golden["newton_cuberoot_adjoint"]
    theta = np.array(n["theta"]); x = np.ones(3)
    dd = np.arange(1, 4)
    for _ in range(60):
        x = x - orc.sparse_solver(dd, dd, 3 * x ** 2, x ** 3 - theta)
    assert np.allclose(x, n["expect_x"], rtol=1e-12)
    # implicit-function adjoint: dL/dtheta = -(dF/dtheta)^T J^{-T} dL/dx with F = x^3 - theta
    _, lam = orc.sparse_solver_grad(np.ones(3), x, dd, dd, 3 * x ** 2)
    assert np.allclose(lam, n["expect_grad"], rtol=1e-12)


def test_poisson2d_generator_and_grad(orc):
    n = 6
    rng = np.random.default_rng(233)
    kext = 1.0 + rng.uniform(size=(n + 2, n + 2))
    cols, val, ncols = orc.poisson2d_matrix(n, kext)
    assert ncols.sum() == len(cols) == 5 * n * n - 4 * n
    rows = np.repeat(np.arange(n * n), ncols)
    A = sp.coo_matrix((val, (rows, cols)), shape=(n * n, n * n)).tocsr()
    assert abs(A - A.T).max() < 1e-14                       # symmetric
    assert np.all(np.linalg.eigvalsh(-A.toarray()) > 0)     # -A is SPD
    # GetPoissonGrad is the derivative of <x, A(k) u> w.r.t. kext (GetPoissonMatrix.h:55-72)
    x = rng.standard_normal(n * n); u = rng.standard_normal((n, n))
    uext = np.zeros((n + 2, n + 2)); uext[1:-1, 1:-1] = u
    gk = orc.poisson2d_grad(n, x, uext).reshape(n + 2, n + 2)

    def L(k):
        c, v, nc = orc.poisson2d_matrix(n, k)
        M = sp.coo_matrix((v, (np.repeat(np.arange(n * n), nc), c)), shape=(n * n, n * n)).tocsr()
        return x @ (M @ u.ravel())

    for (a, b) in ((0, 3), (3, 3), (n + 1, 2), (2, n)):
        e = np.zeros_like(kext); e[a, b] = 1e-6
        assert abs((L(kext + e) - L(kext - e)) / 2e-6 - gk[a, b]) < 1e-7, (a, b)


def test_pcg_port_matches_direct_solve(orc):
    n = 12
    ii, jj, vv = orc.facewise_triplets_2d(n)
    rp, ci, va, _ = orc.coo_to_csr(ii, jj, vv, n * n, n * n)
    assert len(vv) == 8 * n * (n - 1) + 4 * n and len(va) == 5 * n * n - 4 * n
    b = np.random.default_rng(233).uniform(size=n * n)
    x, it, rel = orc.pcg_jacobi(rp, ci, va, b, tol=1e-13, maxit=1000)
    assert it > 0 and rel <= 1e-13
    assert np.allclose(x, orc.sparse_solver(ii, jj, vv, b), rtol=1e-10, atol=1e-12)
    rp3, ci3, va3 = orc.poisson3d_csr(5, 4, 3)
    i3, j3, v3 = orc.facewise_triplets_3d(4, shuffle=False)
    rpf, cif, vaf, _ = orc.coo_to_csr(i3, j3, v3, 64, 64)
    rp4, ci4, va4 = orc.poisson3d_csr(4, 4, 4)
    assert np.array_equal(rpf, rp4) and np.array_equal(cif, ci4) and np.array_equal(vaf, va4)
    assert rp3[-1] == 7 * 60 - 2 * (20 + 12 + 15)


# ======================================================================== structural ops (N2, N4)
def test_get_index_golden(orc, golden):
    g = golden["get_index_bool_mask"]                         # test/sparse.jl:316-321
    sel = np.flatnonzero(g["idof"]) + 1                       # findall(idof)
    I, J, V = orc.sparse_indexing([1, 2], [1, 2], g["diag"], 2, 2, sel, sel)
    assert _dense_from_triplets(I, J, V, 1, 1).tolist() == g["expect_dense"]


def test_structural_oracle_vs_scipy(orc, golden):
    rng = np.random.default_rng(233)
    A = sp.random(10, 10, 0.3, random_state=1, format="coo")
    B = sp.random(3, 3, 0.6, random_state=2, format="coo")
    g = golden["scatter_update_shape"]                        # test/sparse.jl:294-308
    ii, jj = np.array(g["ii"]), np.array(g["jj"])
    I, J, V = orc.sparse_scatter_update(A.row + 1, A.col + 1, A.data, B.row + 1, B.col + 1, B.data, 10, 10, ii, jj)
    C = A.toarray().copy()
    C[np.ix_(ii - 1, jj - 1)] = B.toarray()
    assert np.array_equal(_dense_from_triplets(I, J, V, 10, 10), C)
    # gradient bookkeeping: kept entries first (input order), block entries after
    gout = rng.standard_normal(len(V))
    govv, guvv = orc.sparse_scatter_update_grad(gout, A.row + 1, A.col + 1, B.nnz, ii, jj)
    inside = np.isin(A.row + 1, ii) & np.isin(A.col + 1, jj)
    assert np.all(govv[inside] == 0) and np.array_equal(govv[~inside], gout[: (~inside).sum()])
    assert np.array_equal(guvv, gout[(~inside).sum():])
    z = golden["zero_out_row_shape"]                          # test/sparse.jl:383-389
    oi, ov = orc.zero_out_row(np.stack([A.row, A.col], 1), A.data, z["bd"])
    Z = A.toarray().copy()
    Z[np.array(z["bd"]) - 1, :] = 0
    assert np.array_equal(_dense_from_triplets(oi[:, 0] + 1, oi[:, 1] + 1, ov, 10, 10), Z)
    gz = orc.zero_out_row_grad(np.stack([A.row, A.col], 1), z["bd"], np.arange(1.0, len(ov) + 1))
    assert np.array_equal(gz[gz != 0], np.arange(1.0, len(ov) + 1)) and np.all(gz[np.isin(A.row + 1, z["bd"])] == 0)
    # A[i1, j1] on a 20 x 30 matrix (test/sparse.jl:181-190)
    A2 = sp.random(20, 30, 0.3, random_state=3, format="coo")
    i1, j1 = np.array([17, 3, 9]), np.array([30, 1, 12])
    I, J, V = orc.sparse_indexing(A2.row + 1, A2.col + 1, A2.data, 20, 30, i1, j1)
    assert np.array_equal(_dense_from_triplets(I, J, V, 3, 3), A2.tocsr()[i1 - 1][:, j1 - 1].toarray())
    gi = orc.sparse_indexing_grad(np.ones(len(V)), I, J, A2.row + 1, A2.col + 1, i1, j1)
    assert gi.sum() == len(V) and np.all(gi[~(np.isin(A2.row + 1, i1) & np.isin(A2.col + 1, j1))] == 0)


def test_spgemm_oracle_vs_scipy(orc):
    # test/sparse.jl:141-162: A(10x5)*B(5x20), diag*B, A*diag
    A = sp.random(10, 5, 0.3, random_state=5, format="coo")
    B = sp.random(5, 20, 0.3, random_state=6, format="coo")
    I, J, V = orc.sparse_sparse_mat_mul(A.row, A.col, A.data, B.row, B.col, B.data, 10, 5, 20)
    assert np.allclose(_dense_from_triplets(I, J, V, 10, 20), (A @ B).toarray(), rtol=1e-15, atol=0)
    assert np.all(np.diff(J * 100 + I) > 0), "emission is column-major"
    d = np.array([1.0, 2, 3, 4, 5])
    I, J, V = orc.diag_sparse_mat_mul(d, B.row, B.col, B.data, 5, 20)
    assert np.allclose(_dense_from_triplets(I, J, V, 5, 20), (sp.diags(d) @ B).toarray(), rtol=1e-15, atol=0)
    I, J, V = orc.sparse_diag_mat_mul(A.row, A.col, A.data, d, 10, 5)
    assert np.allclose(_dense_from_triplets(I, J, V, 10, 5), (A @ sp.diags(d)).toarray(), rtol=1e-15, atol=0)


# ======================================================================== Newton-Raphson (N3)
def test_newton_oracle_golden(orc, golden):
    g = golden["newton_cuberoot_adjoint"]                      # test/optim.jl:235-249
    theta = np.array(g["theta"], float)

    def f(th, x):
        return x ** 3 - th, sp.diags(3 * x ** 2)

    x, res, U, conv, it = orc.newton_raphson(f, np.ones(3), theta)
    assert np.allclose(x, g["expect_x"], rtol=1e-12) and conv and it == len(res) == U.shape[1]
    assert res[-1] < 1e-12 <= res[-2]                         # stops one step after the residual passed tol
    grad = orc.newton_adjoint_grad(-sp.identity(3), sp.diags(3 * x ** 2), np.ones(3))
    assert np.allclose(grad, g["expect_grad"], rtol=1e-12)
    r = golden["register_newton_gradient"]                    # test/extra.jl:28-49: y^3 + 1 - θ - x = 0, x = 8, θ = 1
    y, *_ = orc.newton_raphson(lambda th, y: (y ** 3 + 1.0 - th - 8.0, sp.diags(3 * y ** 2)), np.ones(5), np.ones(5))
    assert np.allclose(y, 2.0, rtol=1e-12)
    gth = orc.newton_adjoint_grad(-sp.identity(5), sp.diags(3 * y ** 2), np.ones(5))
    assert np.allclose(gth, r["expect_grad_each"], rtol=1e-12)


def test_trisolve_oracle_vs_dense_and_fd(orc):
    rng = np.random.default_rng(233)                          # test/sparse.jl:352-364
    n = 10
    a, c, b, d = rng.uniform(size=n - 1), rng.uniform(size=n - 1), rng.uniform(size=n) + 10, rng.uniform(size=n)
    A = np.diag(b) + np.diag(a, -1) + np.diag(c, 1)
    x = orc.tri_solve(a, b, c, d)
    assert np.linalg.norm(x - np.linalg.solve(A, d)) < 1e-13
    g = rng.standard_normal(n)                                # loss = <g, x>
    ga, gb, gc, gd = orc.tri_solve_grad(g, x, a, b, c)
    assert np.allclose(gd, np.linalg.solve(A.T, g), rtol=1e-12)
    eps = 1e-6
    for arr, grad, k in ((a, ga, 3), (b, gb, 4), (c, gc, 5), (d, gd, 6)):
        p = arr.copy(); p[k] += eps
        args = {id(a): a, id(b): b, id(c): c, id(d): d}
        args[id(arr)] = p
        fd = (g @ orc.tri_solve(args[id(a)], args[id(b)], args[id(c)], args[id(d)]) - g @ x) / eps
        assert abs(fd - grad[k]) <= 1e-5 * max(1.0, abs(grad[k]))


def test_least_square_oracle_golden_and_fd(orc, golden):
    g = golden["least_squares_adjacent"]                      # test/sparse.jl:131-139
    u = orc.sparse_least_square(g["ii"], g["jj"], g["vv"], g["ff"], g["m"], g["n"])
    assert np.linalg.norm(u - np.array(g["expect"])) < g["tol"]
    rng = np.random.default_rng(233)
    m, n, T = 9, 4, 30
    ii, jj, vv = rng.integers(1, m + 1, T), rng.integers(1, n + 1, T), rng.standard_normal(T)
    ii[:n], jj[:n] = np.arange(1, n + 1), np.arange(1, n + 1)          # full column rank
    vv[:n] += 5.0
    f, w = rng.standard_normal(m), rng.standard_normal(n)
    u = orc.sparse_least_square(ii, jj, vv, f, m, n)
    gvv, gf = orc.sparse_least_square_grad(w, u, ii, jj, vv, f, m, n)  # loss = <w, u>
    eps = 1e-6
    for k in (0, 7, 19):
        v2 = vv.copy(); v2[k] += eps
        fd = (w @ orc.sparse_least_square(ii, jj, v2, f, m, n) - w @ u) / eps
        assert abs(fd - gvv[k]) <= 2e-5 * max(1.0, abs(gvv[k]))
    f2 = f.copy(); f2[3] += eps
    assert abs((w @ orc.sparse_least_square(ii, jj, vv, f2, m, n) - w @ u) / eps - gf[3]) <= 2e-5


# ======================================================================== randomized oracle pinning (hypothesis)
def test_oracle_assembly_and_compress_randomized(orc):
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 12), st.integers(1, 12), st.lists(st.tuples(st.integers(0, 11), st.integers(0, 11),
                                                                     st.floats(-4, 4, allow_nan=False, width=32)), max_size=80))
    def check(m, n, trips):
        trips = [(i % m, j % n, float(v)) for i, j, v in trips]
        ii = np.array([t[0] for t in trips], np.int64) + 1
        jj = np.array([t[1] for t in trips], np.int64) + 1
        vv = np.array([t[2] for t in trips], float)
        rp, ci, va, slot = orc.coo_to_csr(ii, jj, vv, m, n, base=1)
        D = np.zeros((m, n))
        for i, j, v in trips:                      # input-order accumulation, like setFromTriplets
            D[i, j] += v
        R = np.zeros((m, n))
        for r in range(m):
            cols = ci[rp[r]:rp[r + 1]]
            assert np.all(np.diff(cols) > 0)       # sorted, merged
            R[r, cols] = va[rp[r]:rp[r + 1]]
        assert np.array_equal(R, D)
        for t, (i, j, _) in enumerate(trips):      # slot maps every triplet to its stored entry
            e = slot[t]
            assert ci[e] == j and rp[i] <= e < rp[i + 1]
        # compress on the row-major sorted stream
        order = np.lexsort((np.arange(len(trips)), jj, ii)) if trips else np.zeros(0, int)
        idx2 = np.stack([ii[order] - 1, jj[order] - 1], 1) if trips else np.zeros((0, 2), np.int64)
        oi, ov = orc.compress(idx2, vv[order] if trips else vv)
        assert len(ov) == len(va) and (len(ov) == 0 or np.allclose(ov, va, rtol=1e-15, atol=1e-15))
        assert all((ci[rp[r]:rp[r + 1]] == oi[oi[:, 0] == r, 1]).all() for r in range(m)) if len(ov) else True

    check()
