"""Parity tests proper: the CUDA path, called through the C-ABI (ctypes -> libadcme_b200.so),
against the CPU oracle on the same seeded inputs, against the reference's golden vectors, and —
at BASELINE.json's sizes — through size-independent properties.

Bars (BASELINE.json north_star): assembly rowptr/colind bit-exact; fp64 values and SpMV within
1e-12 relative; solve / gradient results within 1e-10 relative residual.
"""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

pytestmark = pytest.mark.gpu

TOL_VAL = 1e-12     # assembly values, SpMV / SpMM, value- and dense-gradients of SpMM
TOL_SOLVE = 1e-10   # solve and adjoint: relative residual / relative error vs the direct oracle


def rel_err(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    d = np.max(np.abs(b)) if b.size else 0.0
    return float(np.max(np.abs(a - b)) / d) if d > 0 else float(np.max(np.abs(a - b))) if a.size else 0.0


def rand_triplets(rng, m, n, T, dup_frac=0.3):
    ii, jj = rng.integers(1, m + 1, T), rng.integers(1, n + 1, T)
    k = int(T * dup_frac)
    src = rng.integers(0, T, k)
    dst = rng.integers(0, T, k)
    ii[dst], jj[dst] = ii[src], jj[src]          # forced duplicates
    return ii.astype(np.int64), jj.astype(np.int64), rng.standard_normal(T)


# ======================================================================== assembly (a1, a4, a5)
@pytest.mark.parametrize("m,n,T", [(1, 1, 1), (6, 5, 6), (37, 29, 2000), (1000, 1000, 50000), (300, 70000, 40000),
                                   (70000, 300, 40000), (5, 5, 0)])
def test_coo_to_csr_bit_exact(pkg, orc, m, n, T):
    rng = np.random.default_rng(233)
    ii, jj, vv = rand_triplets(rng, m, n, T)
    A = pkg.SparseTensor(ii, jj, vv, m, n)
    rp, ci, va = A.csr()
    rp0, ci0, va0, slot0 = orc.coo_to_csr(ii, jj, vv, m, n, base=1)
    assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0)       # bit-exact integers
    assert np.array_equal(va, va0)                                    # same summation order -> bit-exact fp64
    assert np.array_equal(A.slot_map(), slot0)
    assert A.query() == (m, n, len(va0), T)


def test_coo_to_csr_golden_constructor(pkg, golden):
    g = golden["sparse_constructor"]                                  # test/sparse.jl:5-10
    V = np.random.default_rng(233).uniform(size=6)
    A = pkg.SparseTensor(np.array(g["I"]), np.array(g["J"]), V, g["m"], g["n"])
    S = sp.coo_matrix((V, (np.array(g["I"]) - 1, np.array(g["J"]) - 1)), shape=(g["m"], g["n"])).tocsr()
    assert abs(A.to_scipy() - S).max() == 0
    assert A.size() == (6, 5) and A.size(1) == 6 and A.size(2) == 5


def test_duplicate_summation_order_is_input_order(pkg):
    ii, jj = np.array([1, 1, 1, 2]), np.array([1, 1, 1, 2])
    vv = np.array([1e16, 1.0, -1e16, 5.0])
    rp, ci, va = pkg.SparseTensor(ii, jj, vv, 2, 2).csr()
    assert va[0] == (1e16 + 1.0) - 1e16 and va[1] == 5.0


def test_index_range_is_checked(pkg):
    for ii, jj in (([3], [1]), ([1], [3]), ([0], [1])):
        with pytest.raises(pkg.ADCMEError) as e:
            pkg.SparseTensor(np.array(ii), np.array(jj), np.array([1.0]), 2, 2).csr()
        assert e.value.code == -4


def test_base0_int32_and_device_inputs(pkg, orc):
    import torch

    rng = np.random.default_rng(7)
    m, n, T = 200, 150, 5000
    ii, jj, vv = rand_triplets(rng, m, n, T)
    h = C.c_int64(0)
    i32, j32 = (ii - 1).astype(np.int32), (jj - 1).astype(np.int32)
    ti, tj, tv = (torch.as_tensor(x).cuda() for x in (i32, j32, vv))
    pkg.check(pkg.lib.ab_csr_from_coo32(T, ti.data_ptr(), tj.data_ptr(), tv.data_ptr(), m, n, 0, C.byref(h)))
    A = pkg.SparseTensor.from_handle(h.value)
    rp, ci, va = A.csr()
    rp0, ci0, va0, _ = orc.coo_to_csr(ii, jj, vv, m, n, base=1)
    assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0) and np.array_equal(va, va0)


def test_update_values_reuses_pattern(pkg, orc):
    rng = np.random.default_rng(11)
    m = n = 400
    ii, jj, vv = rand_triplets(rng, m, n, 6000)
    A = pkg.SparseTensor(ii, jj, vv, m, n)
    A.csr()
    vv2 = rng.standard_normal(len(vv))
    A._refresh_values(vv2)
    rp, ci, va = A.csr()
    rp0, ci0, va0, _ = orc.coo_to_csr(ii, jj, vv2, m, n)
    assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0) and np.array_equal(va, va0)


def test_facewise_c1_assembly(pkg, orc):
    n = 100                                                           # BASELINE configs[0]
    ii, jj, vv = orc.facewise_triplets_2d(n, seed=233)
    assert len(vv) == 79600
    A = pkg.SparseTensor(ii, jj, vv, n * n, n * n)
    rp, ci, va = A.csr()
    rp0, ci0, va0, _ = orc.coo_to_csr(ii, jj, vv, n * n, n * n)
    assert len(va) == 49600
    assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0) and np.array_equal(va, va0)
    assert set(np.unique(va)) == {-1.0, 4.0}


# ======================================================================== compress / rows (a4, a6)
def test_compress_golden_and_random(pkg, orc, golden):
    g = golden["compress"]                                            # test/sparse.jl:366-381
    A = pkg.SparseTensor(np.array(g["indices_1based"])[:, 0], np.array(g["indices_1based"])[:, 1], np.array(g["v"]), 3, 3)
    Ac = pkg.compress(A)
    assert (np.stack([Ac.ii - 1, Ac.jj - 1], 1)).tolist() == g["expect_indices_0based"]
    assert np.asarray(Ac.vv).tolist() == g["expect_values"]
    rng = np.random.default_rng(233)
    N = 30000
    idx = np.sort(rng.integers(0, 3000, N))
    idx2 = np.stack([idx // 50, idx % 50], 1).astype(np.int64)
    v = rng.standard_normal(N)
    oi, ov = pkg.compress_indices(idx2, v)
    oi0, ov0 = orc.compress(idx2, v)
    assert np.array_equal(oi, oi0) and np.array_equal(ov, ov0)
    gnv = rng.standard_normal(len(ov0))
    assert np.array_equal(pkg.compress_grad(idx2, gnv), orc.compress_grad(idx2, gnv))
    oi, ov = pkg.compress_indices(np.zeros((0, 2), np.int64), np.zeros(0))
    assert oi.shape == (0, 2) and ov.shape == (0,)
    op = pkg.load_op_and_grad(pkg.libadcme, "sparse_compress", multiple=True)
    _, gv = op.grad(gnv, oi0, ov0, idx2, v)
    assert np.array_equal(gv, orc.compress_grad(idx2, gnv))


def test_mpi_create_matrix_golden_and_random(pkg, orc, golden):
    g = golden["mpi_create_matrix"]
    op = pkg.load_system_op("mpi_create_matrix")
    rows, ncols, cols, vals = op(np.array(g["indices"]), np.array(g["values"]), g["ilower"], g["iupper"])
    assert rows.tolist() == g["expect_rows"] and ncols.tolist() == g["expect_ncols"] and cols.tolist() == g["expect_cols"]
    assert rows.dtype == np.int32 and list(vals) == g["expect_values"]
    rng = np.random.default_rng(233)
    idx = np.sort(rng.integers(0, 100000, 20000))
    idx2 = np.stack([idx // 100, idx % 100], 1).astype(np.int64)
    r, nc, c, _ = op(idx2, np.zeros(len(idx)), 77, 0)
    r0, nc0, c0 = orc.coo_to_rows(idx2, 77)
    assert np.array_equal(r, r0) and np.array_equal(nc, nc0) and np.array_equal(c, c0)


# ======================================================================== SparseAssembler (a3)
def test_assembler_golden(pkg, golden):
    for key in ("assembler_tol_filter", "assembler_duplicates", "assembler_docstring_example2"):
        g = golden[key]
        h = pkg.SparseAssembler(g["handle"], g["n"], g["tol"])
        ops = [pkg.accumulate(h, a["row"], a["cols"], a["vals"]) for a in g["adds"]]
        J = pkg.assemble(g["m"], g["ncol"], ops)
        assert np.linalg.norm(J.toarray() - np.array(g["expect_dense"], float)) < 1e-8, key
        assert pkg.lib.ab_asm_nnz(g["handle"]) == -2                  # handle destroyed by the dump


def test_assembler_matches_oracle_order(pkg, orc):
    rng = np.random.default_rng(233)
    nrow, tol = 50, 0.3
    h = pkg.SparseAssembler(4242, nrow, tol)
    acc = orc.Accumulator(nrow, tol)
    for _ in range(300):
        row = int(rng.integers(1, nrow + 1)); k = int(rng.integers(1, 8))
        cols = rng.integers(1, 200, k); vals = rng.standard_normal(k)
        pkg.accumulate(h, row, cols, vals); acc.add(row, cols, vals)
    # bulk device-side add keeps insertion order too
    import torch

    kk = 5000
    br = rng.integers(1, nrow + 1, kk).astype(np.int32); bc = rng.integers(1, 200, kk).astype(np.int32)
    bv = rng.standard_normal(kk)
    tr, tc, tv = (torch.as_tensor(x).cuda() for x in (br, bc, bv))
    pkg.check(pkg.lib.ab_asm_add_coo(h, tr.data_ptr(), tc.data_ptr(), tv.data_ptr(), kk))
    for r_, c_, v_ in zip(br, bc, bv):
        acc.add(int(r_), [c_], [v_])
    pkg.accumulate(h, 3, [9], [9.0]); acc.add(3, [9], [9.0])
    assert pkg.lib.ab_asm_nnz(h) == acc.nnz()
    r, c, v = pkg.load_system_op("sparse_accumulator_copy", False)(h)
    r0, c0, v0 = acc.copy()
    assert np.array_equal(r, r0) and np.array_equal(c, c0) and np.array_equal(v, v0)
    with pytest.raises(pkg.ADCMEError):
        pkg.accumulate(pkg.SparseAssembler(4243, 5, 0.0), 6, [1], [1.0])
    pkg.lib.ab_asm_destroy(4243)


# ======================================================================== SpMV / SpMM (a7, a8)
@pytest.mark.parametrize("m,n,T,k", [(10, 10, 30, 1), (10, 10, 30, 10), (500, 300, 9000, 1), (500, 300, 9000, 3),
                                     (300, 500, 9000, 32), (200, 200, 4000, 70), (2000, 2000, 300000, 1),
                                     (64, 64, 0, 4)])
def test_spmm_parity(pkg, orc, m, n, T, k):
    rng = np.random.default_rng(233)
    ii, jj, vv = rand_triplets(rng, m, n, T)
    A = pkg.SparseTensor(ii, jj, vv, m, n)
    idx2 = np.stack([ii - 1, jj - 1], 1)
    X = rng.uniform(-1, 1, n) if k == 1 else rng.uniform(-1, 1, (n, k))
    Y = A @ X
    Y0 = orc.spmm_coo(idx2, vv, m, X)
    assert Y.shape == Y0.shape and rel_err(Y, Y0) <= TOL_VAL
    # o * s  = (s' o')'  (src/sparse.jl:247-253)
    Z = rng.uniform(-1, 1, m) if k == 1 else rng.uniform(-1, 1, (k, m))
    W = Z @ A
    W0 = orc.spmm_coo(idx2, vv, n, Z if k == 1 else Z.T.copy(), adjoint_a=True)
    W0 = W0 if k == 1 else W0.T
    assert W.shape == W0.shape and rel_err(W, W0) <= TOL_VAL
    # gradients (TF _SparseTensorDenseMatMulGrad)
    G = rng.uniform(-1, 1, Y0.shape)
    if T:
        gv = np.empty(T)
        kx = np.ascontiguousarray(X); kg = np.ascontiguousarray(G)
        pkg.check(pkg.lib.ab_spmm_grad_values(A.handle(), kg.ctypes.data, kx.ctypes.data, k, gv.ctypes.data))
        assert rel_err(gv, orc.spmm_grad_values(idx2, G, X)) <= TOL_VAL
    gx = np.empty_like(np.ascontiguousarray(X))
    kg = np.ascontiguousarray(G)
    pkg.check(pkg.lib.ab_spmm_grad_dense(A.handle(), kg.ctypes.data, k, gx.ctypes.data))
    assert rel_err(gx, orc.spmm_coo(idx2, vv, n, G, adjoint_a=True)) <= TOL_VAL


@pytest.mark.parametrize("k", [32, 40, 128, 150])
def test_spmm_power_law_long_rows(pkg, orc, k):
    """C5-shaped: a few rows far longer than LONG_ROW (4096) are cut into chunks and recombined."""
    rng = np.random.default_rng(233)
    m = n = 3000
    deg = np.minimum(np.floor(rng.uniform(1e-9, 1, m) ** (-1 / 1.1)), 20000).astype(np.int64)
    deg[5], deg[6], deg[2999] = 20000, 4097, 9000
    ii = np.repeat(np.arange(1, m + 1), deg)
    jj = rng.integers(1, n + 1, ii.size)
    vv = rng.standard_normal(ii.size)
    A = pkg.SparseTensor(ii, jj, vv, m, n)
    idx2 = np.stack([ii - 1, jj - 1], 1)
    X = rng.uniform(-1, 1, (n, k))
    assert rel_err(A @ X, orc.spmm_coo(idx2, vv, m, X)) <= TOL_VAL
    G = rng.uniform(-1, 1, (m, k))
    gv = np.empty(ii.size)
    pkg.check(pkg.lib.ab_spmm_grad_values(A.handle(), G.ctypes.data, X.ctypes.data, k, gv.ctypes.data))
    assert rel_err(gv, orc.spmm_grad_values(idx2, G, X)) <= TOL_VAL
    gx = np.empty_like(X)
    pkg.check(pkg.lib.ab_spmm_grad_dense(A.handle(), G.ctypes.data, k, gx.ctypes.data))
    assert rel_err(gx, orc.spmm_coo(idx2, vv, n, G, adjoint_a=True)) <= TOL_VAL


def test_refinement_reaches_true_residual(pkg):
    """Ill-conditioned SPD system (2-D variable-coefficient operator, 700^2): the recurrence residual
    alone stalls above 1e-10; the refinement step must deliver the TRUE residual (north_star bar)."""
    import torch

    n = 700
    xs = torch.arange(0, n + 2, dtype=torch.float64, device="cuda") / (n + 1)
    X, Y = torch.meshgrid(xs, xs, indexing="ij")
    kext = (1 + 50 * X ** 2 + Y ** 2).contiguous()
    h = C.c_int64(0)
    pkg.check(pkg.lib.ab_poisson2d_csr(n, kext.data_ptr(), -1, C.byref(h)))
    A = pkg.SparseTensor.from_handle(h.value)
    f = torch.rand(n * n, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(233))
    u = torch.empty_like(f)
    it, rr = C.c_int(0), C.c_double(0)
    for method in (b"CG", b"BiCGSTAB"):
        pkg.check(pkg.lib.ab_solve(A.handle(), method, f.data_ptr(), u.data_ptr(), 1e-10, 100000, C.byref(it), C.byref(rr)))
        res = float((A @ u - f).norm() / f.norm())
        assert res <= TOL_SOLVE and abs(rr.value - res) <= 1e-3 * res + 1e-14, (method, res, rr.value)


def test_spmv_all_kernel_variants(pkg, orc):
    """tile kernel (U=8, U=4, R=256/128/64) and the vector kernel give the same SpMV."""
    rng = np.random.default_rng(233)
    cases = []
    n = 48
    cases.append(orc.facewise_triplets_3d(n, shuffle=True) + (n ** 3,))                # 7 nnz/row -> R=256
    m = 3000
    cases.append(rand_triplets(rng, m, m, 12 * m, 0.0) + (m,))                           # ~12/row -> R=128
    cases.append(rand_triplets(rng, m, m, 28 * m, 0.0) + (m,))                           # ~28/row -> R=64
    cases.append(rand_triplets(rng, 500, 500, 120 * 500, 0.0) + (500,))                  # dense rows -> vector kernel
    ii = np.concatenate([np.full(30000, 7), np.arange(1, 501)]); jj = np.concatenate([rng.integers(1, 501, 30000), np.arange(1, 501)])
    cases.append((ii.astype(np.int64), jj.astype(np.int64), rng.standard_normal(len(ii)), 500))   # one very long row
    for ii, jj, vv, N in cases:
        A = pkg.SparseTensor(ii, jj, vv, N, N)
        x = rng.uniform(-1, 1, N)
        y0 = orc.spmm_coo(np.stack([ii - 1, jj - 1], 1), vv, N, x)
        for variant in (-1, 0, 1, 2):
            for off8 in (1.0, 0.0):           # offset-dictionary form of colind on / off
                pkg.check(pkg.lib.ab_set_option(b"spmv_variant", float(variant)))
                pkg.check(pkg.lib.ab_set_option(b"spmv_off8", off8))
                try:
                    assert rel_err(A @ x, y0) <= TOL_VAL, (N, variant, off8)
                finally:
                    pkg.check(pkg.lib.ab_set_option(b"spmv_variant", -1.0))
                    pkg.check(pkg.lib.ab_set_option(b"spmv_off8", 1.0))


def test_spmv_torch_device_tensors_and_autograd(pkg, orc):
    import torch

    rng = np.random.default_rng(233)
    m, n, T, k = 120, 90, 1500, 4
    ii, jj, vv = rand_triplets(rng, m, n, T)
    idx2 = np.stack([ii - 1, jj - 1], 1)
    v_t = torch.tensor(vv, device="cuda", requires_grad=True)
    X_t = torch.tensor(rng.uniform(-1, 1, (n, k)), device="cuda", requires_grad=True)
    A = pkg.SparseTensor(ii, jj, v_t, m, n)
    Y = A @ X_t
    assert Y.is_cuda and rel_err(Y.detach().cpu().numpy(), orc.spmm_coo(idx2, vv, m, X_t.detach().cpu().numpy())) <= TOL_VAL
    G = torch.tensor(rng.uniform(-1, 1, (m, k)), device="cuda")
    (Y * G).sum().backward()
    assert rel_err(v_t.grad.cpu().numpy(), orc.spmm_grad_values(idx2, G.cpu().numpy(), X_t.detach().cpu().numpy())) <= TOL_VAL
    assert rel_err(X_t.grad.cpu().numpy(), orc.spmm_coo(idx2, vv, n, G.cpu().numpy(), adjoint_a=True)) <= TOL_VAL


def test_adjoint_matches_reference_semantics(pkg):
    rng = np.random.default_rng(233)                                   # test/sparse.jl:36-40
    S = sp.random(10, 5, 0.3, random_state=233, format="csr")
    A = pkg.SparseTensor(S)
    assert abs(A.T.to_scipy() - S.T.tocsr()).max() == 0
    h = pkg.SparseTensor.from_handle(A.handle())
    At = h.adjoint()
    assert abs(At.to_scipy() - S.T.tocsr()).max() == 0
    h._h = None


# ======================================================================== solve + adjoint (a9, a10, a11)
def spd_case(orc, n):
    ii, jj, vv = orc.facewise_triplets_2d(n, seed=233)
    return ii, jj, vv, n * n


def test_solve_c1_poisson_all_methods(pkg, orc):
    ii, jj, vv, N = spd_case(orc, 100)                                 # BASELINE configs[0]
    rp, ci, va, _ = orc.coo_to_csr(ii, jj, vv, N, N)
    rng = np.random.default_rng(233)
    ustar = rng.uniform(0, 1, N)
    f = orc.csr_spmm(rp, ci, va, ustar)
    u0 = orc.sparse_solver(ii, jj, vv, f)
    A = pkg.SparseTensor(ii, jj, vv, N, N)
    for method in ("CG", "BiCGSTAB", "SparseLU", "SimplicialLLT", "SimplicialLDLT", "SparseQR", "auto"):
        u = pkg.backslash(A, f, method)
        res = np.linalg.norm(orc.csr_spmm(rp, ci, va, u) - f) / np.linalg.norm(f)
        assert res <= TOL_SOLVE, (method, res)
        assert np.linalg.norm(u - u0) / np.linalg.norm(u0) <= 1e-8, method      # cond ~ 6e3
    with pytest.raises(pkg.ADCMEError):
        pkg.backslash(A, f, "Cholmod")
    # op-level call with the reference signature
    u = pkg.load_system_op("sparse_solver")(ii, jj, vv, f, "SparseLU")
    assert np.linalg.norm(orc.csr_spmm(rp, ci, va, u) - f) / np.linalg.norm(f) <= TOL_SOLVE


def test_cg_scaled_form_matches_classic_pcg(pkg, orc):
    """CG runs on the symmetrically scaled system S A S (11 vector passes); it must agree with the
    classic Jacobi-PCG recurrence and with the direct oracle on a strongly varying diagonal."""
    rng = np.random.default_rng(233)
    n = 40
    kext = np.exp(rng.uniform(-2, 2, (n + 2, n + 2)))        # diagonal spans ~e^4
    h = C.c_int64(0)
    pkg.check(pkg.lib.ab_poisson2d_csr(n, kext.ctypes.data, -1, C.byref(h)))
    A = pkg.SparseTensor.from_handle(h.value)
    S = A.to_scipy().tocoo()
    f = rng.uniform(size=n * n)
    u0 = orc.sparse_solver(S.row + 1, S.col + 1, S.data, f)
    sols = []
    for scaled in (1.0, 0.0):
        pkg.check(pkg.lib.ab_set_option(b"cg_scaled", scaled))
        try:
            u = pkg.backslash(A, f, "CG")
        finally:
            pkg.check(pkg.lib.ab_set_option(b"cg_scaled", 1.0))
        assert np.linalg.norm(S @ u - f) / np.linalg.norm(f) <= TOL_SOLVE, scaled
        assert np.linalg.norm(u - u0) / np.linalg.norm(u0) <= 1e-8, scaled
        sols.append(u)
    assert np.linalg.norm(sols[0] - sols[1]) / np.linalg.norm(u0) <= 1e-8
    # a non-positive diagonal entry cannot be scaled (sqrt): the classic recurrence takes over
    ii = np.array([1, 2, 3]); vv = np.array([1.0, -2.0, 3.0])
    u = pkg.backslash(pkg.SparseTensor(ii, ii, vv, 3, 3), np.ones(3), "CG")
    assert np.allclose(u, [1.0, -0.5, 1.0 / 3.0], rtol=1e-12)


def test_legacy_method_strings_try_cg_on_symmetric_matrices(pkg, orc):
    """SparseLU / SparseQR / auto: CG first when A == A' bitwise with a positive diagonal, BiCGSTAB (+ dense
    LU) otherwise or when CG breaks down — the answer is vetted by the true residual either way."""
    rng = np.random.default_rng(233)
    # symmetric INDEFINITE with a positive diagonal: CG must give up and the general path must answer
    n = 60
    B = sp.random(n, n, 0.1, random_state=5, format="csr")
    S = (B + B.T + sp.identity(n) * 0.05).tocoo()
    assert np.linalg.eigvalsh(S.toarray()).min() < 0 and S.diagonal().min() > 0
    A = pkg.SparseTensor(S.row + 1, S.col + 1, S.data, n, n)
    f = rng.uniform(size=n)
    u = pkg.backslash(A, f, "SparseLU")
    assert np.linalg.norm(S @ u - f) / np.linalg.norm(f) <= TOL_SOLVE
    # SPD through the legacy strings: same answer as the explicit CG, and the adjoint needs no transpose
    ii, jj, vv, N = spd_case(orc, 30)
    P = pkg.SparseTensor(ii, jj, vv, N, N)
    f = rng.uniform(size=N)
    u1, u2 = pkg.backslash(P, f, "SparseLU"), pkg.backslash(P, f, "CG")
    assert np.array_equal(u1, u2)
    g = rng.uniform(size=N)
    op = pkg.load_system_op("sparse_solver")
    r1 = op.grad(g, u1, ii, jj, vv, f, "SparseLU")
    r2 = op.grad(g, u1, ii, jj, vv, f, "SimplicialLLT")
    assert np.array_equal(r1[2], r2[2]) and np.array_equal(r1[3], r2[3])
    pkg.check(pkg.lib.ab_set_option(b"auto_cg", 0.0))          # the switch: plain BiCGSTAB again
    try:
        u3 = pkg.backslash(P, f, "SparseLU")
    finally:
        pkg.check(pkg.lib.ab_set_option(b"auto_cg", 1.0))
    assert np.linalg.norm(u3 - u1) <= 1e-9 * np.linalg.norm(u1) and not np.array_equal(u3, u1)


def test_solve_random_nonsymmetric_like_reference_test(pkg, orc):
    rng = np.random.default_rng(233)                                   # test/sparse.jl:71-78
    # (10, 0.1) is the reference's own case; the larger ones exercise BiCGSTAB on diagonally dominant
    # matrices and, for the indefinite 200 x 200 one, the small dense-LU fallback of "SparseLU"
    for d, dens, shift in ((10, 0.1, 1.0), (200, 0.02, 1.0), (3000, 0.002, 4.0), (40000, 0.0002, 6.0)):
        S = (sp.identity(d) * shift + sp.random(d, d, dens, random_state=233)).tocoo()
        ii, jj, vv = S.row + 1, S.col + 1, S.data
        b = rng.uniform(size=d)
        u = pkg.backslash(pkg.SparseTensor(ii, jj, vv, d, d), b)
        assert np.linalg.norm(S @ u - b) / np.linalg.norm(b) <= TOL_SOLVE, d
        if d <= 3000:       # a direct LU of an unstructured 40000^2 matrix fills in: residual check only
            u0 = orc.sparse_solver(ii, jj, vv, b)
            assert np.linalg.norm(u - u0) / np.linalg.norm(u0) <= 1e-8, d
    # test/sparse.jl:323-333 factorises a dense-ish random 10 x 10 (indefinite): legacy strings must cope
    S = sp.random(10, 10, 0.7, random_state=233).tocoo()
    ii, jj, vv = S.row + 1, S.col + 1, S.data
    b = rng.uniform(size=10)
    u = pkg.backslash(pkg.SparseTensor(ii, jj, vv, 10, 10), b, "SparseLU")
    assert np.linalg.norm(u - np.linalg.solve(S.toarray(), b)) / np.linalg.norm(u) <= 1e-9


def test_solve_indefinite_saddle_point_beyond_2048_uses_the_wide_dense_fallback(pkg):
    """A saddle-point system [[K, B'], [B, 0]] with 3000 unknowns: zero diagonal block, indefinite, unsymmetric after the
    perturbation — with the iteration budget cut to 5 no Jacobi-Krylov method gets anywhere, the reference's SparseLU
    would not care.  The legacy strings must cope too: above the one-CTA dense LU (n <= 2048) the same partial-pivot
    elimination runs over the whole GPU (n <= "dense_lu_max", default 8192).  Explicit "BiCGSTAB" still reports the
    failure instead of a wrong answer."""
    rng = np.random.default_rng(233)
    n1, n2 = 2000, 1000
    K = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(n1, n1)) + 0.01 * sp.random(n1, n1, 0.001, random_state=1)
    B = sp.random(n2, n1, 0.004, random_state=2) + sp.eye(n2, n1)
    S = sp.bmat([[K, B.T], [B, None]]).tocoo()
    d = n1 + n2
    ii, jj, vv = S.row + 1, S.col + 1, S.data
    b = rng.uniform(size=d)
    A = pkg.SparseTensor(ii, jj, vv, d, d)
    u0 = spla.spsolve(S.tocsc(), b)
    pkg.check(pkg.lib.ab_set_option(b"maxit", 5.0))
    try:
        u = pkg.backslash(A, b, "SparseLU")
        assert np.linalg.norm(S @ u - b) / np.linalg.norm(b) <= TOL_SOLVE
        assert np.linalg.norm(u - u0) / np.linalg.norm(u0) <= 1e-8
        with pytest.raises(pkg.ADCMEError):
            pkg.backslash(A, b, "BiCGSTAB")
        # and the cliff itself moved, it did not vanish: above "dense_lu_max" the failure is reported
        pkg.check(pkg.lib.ab_set_option(b"dense_lu_max", 2048.0))
        with pytest.raises(pkg.ADCMEError):
            pkg.backslash(A, b, "SparseLU")
    finally:
        pkg.check(pkg.lib.ab_set_option(b"dense_lu_max", 8192.0))
        pkg.check(pkg.lib.ab_set_option(b"maxit", 0.0))


def test_solve_gradients_parity_and_fd(pkg, orc):
    rng = np.random.default_rng(233)
    ii, jj, vv, N = spd_case(orc, 30)
    vv = vv * rng.uniform(0.9, 1.1, len(vv))                            # unsymmetric perturbation ...
    ii = np.concatenate([ii, np.arange(1, N + 1)]); jj = np.concatenate([jj, np.arange(1, N + 1)])
    vv = np.concatenate([vv, np.ones(N)])                               # ... plus a mass term: well conditioned
    f = rng.uniform(size=N)
    g = rng.uniform(-1, 1, N)
    op = pkg.load_system_op("sparse_solver")
    u = op(ii, jj, vv, f, "SparseLU")
    u0 = orc.sparse_solver(ii, jj, vv, f)
    assert np.linalg.norm(u - u0) / np.linalg.norm(u0) <= 1e-9
    gi, gj, gvv, gf, gm = op.grad(g, u, ii, jj, vv, f, "SparseLU")
    assert gi is None and gj is None and gm is None
    gvv0, gf0 = orc.sparse_solver_grad(g, u0, ii, jj, vv)
    assert np.linalg.norm(gf - gf0) / np.linalg.norm(gf0) <= 1e-9
    assert np.linalg.norm(gvv - gvv0) / np.linalg.norm(gvv0) <= 1e-9
    # FD-vs-AD convergence order in the style of src/kit.jl:60-101 / gradtest.jl
    L = lambda v: float(g @ op(ii, jj, v, f, "SparseLU"))
    dv = rng.standard_normal(len(vv))
    L0, slope = L(vv), float(gvv @ dv)
    e1 = [abs(L(vv + t * dv) - L0) for t in (1e-2, 1e-3, 1e-4)]
    e2 = [abs(L(vv + t * dv) - L0 - t * slope) for t in (1e-2, 1e-3, 1e-4)]
    assert all(b < a for a, b in zip(e1, e2)), (e1, e2, slope)
    assert e2[1] < e2[0] * 0.02 and e2[2] < e2[1] * 0.05, (e1, e2, slope)   # second-order decay


def test_solve_autograd_through_backslash(pkg, orc):
    import torch

    rng = np.random.default_rng(233)                                   # docs/src/tu_sparse.md:104-116
    ii, jj, vv, N = spd_case(orc, 20)
    v_t = torch.tensor(vv, device="cuda", requires_grad=True)
    f_t = torch.tensor(rng.uniform(size=N), device="cuda", requires_grad=True)
    A = pkg.SparseTensor(ii, jj, v_t, N, N)
    sol = pkg.backslash(A, f_t, "SimplicialLLT")
    sol.sum().backward()
    u0 = orc.sparse_solver(ii, jj, vv, f_t.detach().cpu().numpy())
    gvv0, gf0 = orc.sparse_solver_grad(np.ones(N), u0, ii, jj, vv)
    assert np.linalg.norm(v_t.grad.cpu().numpy() - gvv0) / np.linalg.norm(gvv0) <= 1e-9
    assert np.linalg.norm(f_t.grad.cpu().numpy() - gf0) / np.linalg.norm(gf0) <= 1e-9


def test_singular_matrix_raises(pkg, golden):
    n = golden["singular_solve_must_throw"]["n"]                       # test/sparse.jl:335-339
    A = pkg.SparseTensor(np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0), n, n)
    with pytest.raises(pkg.ADCMEError):
        pkg.backslash(A, np.random.default_rng(0).uniform(size=n))
    for method in ("CG", "BiCGSTAB"):
        A = pkg.SparseTensor(np.array([1, 2]), np.array([1, 2]), np.array([1.0, 0.0]), 2, 2)
        with pytest.raises(pkg.ADCMEError):
            pkg.backslash(A, np.array([1.0, 1.0]), method)
    u = pkg.backslash(pkg.SparseTensor(np.array([1, 2]), np.array([1, 2]), np.array([2.0, 4.0]), 2, 2), np.zeros(2))
    assert np.array_equal(u, np.zeros(2))                              # b = 0 -> x = 0


def test_newton_cuberoot_known_answer(pkg, golden):
    g = golden["newton_cuberoot_adjoint"]                              # test/optim.jl:235-249
    theta = np.array(g["theta"]); x = np.ones(3)
    dd = np.arange(1, 4)
    op = pkg.load_system_op("sparse_solver")
    for _ in range(60):
        x = x - op(dd, dd, 3 * x ** 2, x ** 3 - theta, "SparseLU")
    assert np.allclose(x, g["expect_x"], rtol=1e-12)
    _, _, _, lam, _ = op.grad(np.ones(3), x, dd, dd, 3 * x ** 2, np.zeros(3), "SparseLU")
    assert np.allclose(lam, g["expect_grad"], rtol=1e-10)
    g2 = golden["register_newton_gradient"]                            # test/extra.jl:28-49
    y = np.ones(5); d5 = np.arange(1, 6)
    for _ in range(60):
        y = y - op(d5, d5, 3 * y ** 2, y ** 3 + 1.0 - 1.0 - 8.0, "SparseLU")
    assert np.allclose(y, g2["expect_y"], rtol=1e-12)
    _, _, _, lam, _ = op.grad(np.ones(5), y, d5, d5, 3 * y ** 2, np.zeros(5), "SparseLU")
    assert np.allclose(lam, g2["expect_grad_each"], rtol=1e-10)


def test_factorize_and_solve(pkg, orc):
    rng = np.random.default_rng(233)                                   # test/sparse.jl:323-333
    d = 60
    S = (sp.identity(d) * 3 + sp.random(d, d, 0.2, random_state=233)).tocoo()
    ii, jj, vv = S.row + 1, S.col + 1, S.data
    Afac = pkg.factorize(pkg.SparseTensor(ii, jj, vv, d, d))
    for _ in range(2):
        rhs = rng.uniform(size=d)
        v1 = pkg.backslash(Afac, rhs)
        assert np.linalg.norm(v1 - orc.sparse_solver(ii, jj, vv, rhs)) < 1e-8
    fi, fj, fv = pkg.find(Afac[0])
    gout = rng.uniform(-1, 1, d)
    out = pkg.solve(Afac, rhs)
    grhs, _, _, gvv, _ = pkg.load_system_op("solve").grad(gout, out, rhs, fi, fj, fv, Afac[1])
    gvv0, gf0 = orc.sparse_solver_grad(gout, out, fi, fj, fv)
    assert np.linalg.norm(grhs - gf0) / np.linalg.norm(gf0) <= 1e-9
    assert np.linalg.norm(gvv - gvv0) / np.linalg.norm(gvv0) <= 1e-9
    # LRU eviction: max_cache_size 1 drops the older id (lru_cache.cpp:62-106)
    a1 = pkg.factorize(pkg.SparseTensor(ii, jj, vv, d, d), 1)
    a2 = pkg.factorize(pkg.SparseTensor(ii, jj, vv, d, d), 1)
    pkg.solve(a2, rhs)
    with pytest.raises(pkg.ADCMEError):
        pkg.solve(a1, rhs)
    pkg.factorize(pkg.SparseTensor(ii, jj, vv, d, d), 999999)


# ======================================================================== generators (N1)
def test_poisson3d_generator_matches_assembly(pkg, orc):
    n = 12
    h = C.c_int64(0)
    pkg.check(pkg.lib.ab_poisson3d_csr(n, n, n, 0, n ** 3, C.byref(h)))
    G = pkg.SparseTensor.from_handle(h.value)
    rp, ci, va = G.csr()
    ii, jj, vv = orc.facewise_triplets_3d(n, seed=233)
    rp1, ci1, va1 = pkg.SparseTensor(ii, jj, vv, n ** 3, n ** 3).csr()
    rp0, ci0, va0 = orc.poisson3d_csr(n, n, n)
    assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0) and np.array_equal(va, va0)
    assert np.array_equal(rp1, rp0) and np.array_equal(ci1, ci0) and np.array_equal(va1, va0)
    # a slab keeps global column ids
    lo, hi = 3 * n * n, 7 * n * n
    pkg.check(pkg.lib.ab_poisson3d_csr(n, n, n, lo, hi, C.byref(h)))
    rps, cis, vas = pkg.SparseTensor.from_handle(h.value).csr()
    assert np.array_equal(rps, rp0[lo:hi + 1] - rp0[lo]) and np.array_equal(cis, ci0[rp0[lo]:rp0[hi]])


def test_poisson2d_generator_and_grad(pkg, orc):
    n = 33
    rng = np.random.default_rng(233)
    kext = 1.0 + rng.uniform(size=(n + 2, n + 2))
    cols, val, ncols = orc.poisson2d_matrix(n, kext)                   # GetPoissonMatrix.h:15-51
    rows = np.repeat(np.arange(n * n), ncols)
    R = sp.coo_matrix((val, (rows, cols)), shape=(n * n, n * n)).tocsr(); R.sort_indices()
    for sign in (1, -1):
        h = C.c_int64(0)
        pkg.check(pkg.lib.ab_poisson2d_csr(n, kext.ctypes.data, sign, C.byref(h)))
        G = pkg.SparseTensor.from_handle(h.value)
        rp, ci, va = G.csr()
        assert np.array_equal(rp, R.indptr) and np.array_equal(ci, R.indices)
        assert np.array_equal(va, sign * R.data)                       # same arithmetic order -> bit-exact
        x = rng.standard_normal(n * n); uext = np.zeros((n + 2, n + 2)); uext[1:-1, 1:-1] = rng.standard_normal((n, n))
        gk = np.empty((n + 2) * (n + 2))
        pkg.check(pkg.lib.ab_poisson2d_grad(n, x.ctypes.data, uext.ctypes.data, sign, gk.ctypes.data))
        assert rel_err(gk, sign * orc.poisson2d_grad(n, x, uext)) <= 1e-13   # GetPoissonMatrix.h:55-72
        k2 = 1.0 + rng.uniform(size=(n + 2, n + 2))
        pkg.check(pkg.lib.ab_poisson2d_update(h.value, n, k2.ctypes.data, sign))
        c2, v2, nc2 = orc.poisson2d_matrix(n, k2)
        R2 = sp.coo_matrix((v2, (rows, c2)), shape=(n * n, n * n)).tocsr(); R2.sort_indices()
        assert np.array_equal(G.csr()[2], sign * R2.data)
    # inverse-conductivity slice (BASELINE configs[3] shape, small): solve with -A (SPD) by CG
    f = rng.uniform(size=n * n)
    u = pkg.backslash(G, f, "CG")
    assert np.linalg.norm((-R2) @ u - f) / np.linalg.norm(f) <= TOL_SOLVE


# ======================================================================== BASELINE-size properties
def test_c2_spmv_256cubed_properties(pkg):
    """configs[1]: 3-D 7-pt Laplacian 256^3 — properties that need no CPU oracle."""
    import torch

    n = 256
    N = n ** 3
    h = C.c_int64(0)
    pkg.check(pkg.lib.ab_poisson3d_csr(n, n, n, 0, N, C.byref(h)))
    A = pkg.SparseTensor.from_handle(h.value)
    assert A.query() == (N, N, 7 * N - 6 * n * n, 7 * N - 6 * n * n)
    g = torch.Generator(device="cuda").manual_seed(233)
    x = torch.rand(N, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    z = torch.rand(N, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    one = torch.ones(N, dtype=torch.float64, device="cuda")
    Ax, Az = A @ x, A @ z
    # row sums: interior rows sum to 0, every boundary face adds 1
    r = (A @ one).view(n, n, n)
    idx = torch.arange(n, device="cuda")
    edge = ((idx == 0) | (idx == n - 1)).double()
    expect = edge.view(n, 1, 1) + edge.view(1, n, 1) + edge.view(1, 1, n)
    assert torch.equal(r, expect)
    # linearity and symmetry
    lin = A @ (2.0 * x + z)
    assert float((lin - (2.0 * Ax + Az)).abs().max()) <= 1e-12 * float(lin.abs().max())
    assert abs(float(torch.dot(z, Ax)) - float(torch.dot(x, Az))) <= 1e-11 * abs(float(torch.dot(z, Ax)))
    # against the matrix-free stencil evaluated by torch (independent of CSR)
    X = x.view(n, n, n)
    S = 6.0 * X
    S[1:] -= X[:-1]; S[:-1] -= X[1:]
    S[:, 1:] -= X[:, :-1]; S[:, :-1] -= X[:, 1:]
    S[:, :, 1:] -= X[:, :, :-1]; S[:, :, :-1] -= X[:, :, 1:]
    assert float((Ax.view(n, n, n) - S).abs().max()) <= 1e-12 * float(S.abs().max())
    # value gradient (per stored entry) and dense gradient at full size
    dv = torch.empty(A.query()[2], dtype=torch.float64, device="cuda")
    pkg.check(pkg.lib.ab_spmm_grad_values(A.handle(), z.data_ptr(), x.data_ptr(), 1, dv.data_ptr()))
    rp = torch.empty(N + 1, dtype=torch.int32, device="cuda"); ci = torch.empty(A.query()[2], dtype=torch.int32, device="cuda")
    pkg.check(pkg.lib.ab_csr_export(A.handle(), rp.data_ptr(), ci.data_ptr(), None))
    rows = torch.repeat_interleave(torch.arange(N, device="cuda"), (rp[1:] - rp[:-1]).long())
    assert torch.equal(dv, z[rows] * x[ci.long()])
    dx = torch.empty(N, dtype=torch.float64, device="cuda")
    pkg.check(pkg.lib.ab_spmm_grad_dense(A.handle(), z.data_ptr(), 1, dx.data_ptr()))
    assert float((dx - Az).abs().max()) <= 1e-12 * float(Az.abs().max())     # A symmetric: A' z = A z


def test_c2_assembly_large_facewise(pkg, orc):
    """configs[1] assembly at a size the oracle still finishes in seconds (96^3, 10.6 M triplets),
    bit-exact, plus the closed-form structure check."""
    n = 96
    ii, jj, vv = orc.facewise_triplets_3d(n, seed=233)
    A = pkg.SparseTensor(ii, jj, vv, n ** 3, n ** 3)
    rp, ci, va = A.csr()
    rp0, ci0, va0 = orc.poisson3d_csr(n, n, n)
    assert np.array_equal(rp, rp0) and np.array_equal(ci, ci0) and np.array_equal(va, va0)


def test_c3_cg_solve_converges_128cubed(pkg):
    """configs[2] at 128^3: full Jacobi-PCG solve to 1e-10, checked by the true residual."""
    import torch

    n = 128
    N = n ** 3
    h = C.c_int64(0)
    pkg.check(pkg.lib.ab_poisson3d_csr(n, n, n, 0, N, C.byref(h)))
    A = pkg.SparseTensor.from_handle(h.value)
    ustar = torch.rand(N, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(233))
    f = A @ ustar
    u = torch.empty_like(f)
    it, rr = C.c_int(0), C.c_double(0)
    pkg.check(pkg.lib.ab_solve(A.handle(), b"CG", f.data_ptr(), u.data_ptr(), 1e-11, 5000, C.byref(it), C.byref(rr)))
    res = float((A @ u - f).norm() / f.norm())
    assert res <= TOL_SOLVE and 0 < it.value < 2000 and rr.value <= 1e-11
    assert float((u - ustar).norm() / ustar.norm()) <= 1e-7
    # non-convergence is an error, not a silent result
    rc = pkg.lib.ab_solve(A.handle(), b"CG", f.data_ptr(), u.data_ptr(), 1e-11, 5, C.byref(it), C.byref(rr))
    assert rc == -5 and it.value == 5
    # fixed-count entry used by bench.py is deterministic
    u2 = torch.empty_like(f)
    pkg.check(pkg.lib.ab_cg_fixed(A.handle(), f.data_ptr(), u.data_ptr(), 25, C.byref(rr)))
    pkg.check(pkg.lib.ab_cg_fixed(A.handle(), f.data_ptr(), u2.data_ptr(), 25, C.byref(rr)))
    assert torch.equal(u, u2)


@pytest.mark.parametrize("n", [150, 120])
def test_solver_spmv_forms_on_unaligned_grids(pkg, n):
    """Grids whose plane stride n^2 is no multiple of the 1024-row tile (150^2 = 22500: not even of 16; 120^2 = 14400 =
    16 * 900): every SpMV form the solver can take — whichever kernel the form falls back to on this grid — reproduces the
    plain int32 kernel bit for bit.  (A marching ring with a stride rounded to the tile size once gave wrong products
    exactly here.)"""
    import torch

    keys = (b"spmv_rowpat", b"spmv_code", b"spmv_off8", b"spmv_rowwin", b"spmv_rowmarch", b"spmv_zmarch")
    forms = (("zmarch", 1, 1, 1, 1, 1, 1), ("rowmarch", 1, 1, 1, 1, 1, 0), ("rowwin", 1, 1, 1, 1, 0, 0), ("rowpat", 1, 1, 1, 0, 0, 0),
             ("code", 0, 1, 1, 0, 0, 0), ("off8", 0, 0, 1, 0, 0, 0), ("plain", 0, 0, 0, 0, 0, 0))
    N = n ** 3
    h = C.c_int64(0)
    pkg.check(pkg.lib.ab_poisson3d_csr(n, n, n, 0, N, C.byref(h)))
    A = pkg.SparseTensor.from_handle(h.value)
    gen = torch.Generator(device="cuda").manual_seed(233)
    x = torch.randn(N, dtype=torch.float64, device="cuda", generator=gen)
    ys = {}
    try:
        for name, *vals in forms:
            for k, v in zip(keys, vals):
                pkg.check(pkg.lib.ab_set_option(k, float(v)))
            y = torch.empty_like(x)
            pkg.check(pkg.lib.ab_spmv_solver_form(A.handle(), x.data_ptr(), y.data_ptr()))
            ys[name] = y
    finally:
        for k in keys:
            pkg.check(pkg.lib.ab_set_option(k, 1.0))
    for name, y in ys.items():
        assert torch.equal(y, ys["plain"]), (n, name)
    S = A.to_scipy()
    dinv = 1.0 / np.sqrt(S.diagonal())
    y0 = dinv * (S @ (dinv * x.cpu().numpy()))
    assert np.max(np.abs(ys["plain"].cpu().numpy() - y0)) <= 1e-13 * np.max(np.abs(y0))
    zi = C.c_double(0)
    pkg.check(pkg.lib.ab_csr_get_info(A.handle(), b"zmarch", C.byref(zi)))
    assert zi.value == (1.0 if n == 120 else 0.0)


@pytest.mark.parametrize("n", [1040, 2048])
def test_solver_spmv_forms_on_2d_grids_with_long_lines(pkg, n):
    """2-D 5-point stencil with lines of >= 1024 points: the line length is the marching stride (1040 = 16 * 65: strips
    that do not coincide with tiles; 2048: two tiles per line), the dominant pattern has five entries — the
    spmv_zmarch_kernel<G, 5, 2> instantiation.  Every form against the plain int32 kernel, bit for bit."""
    import torch

    keys = (b"spmv_rowpat", b"spmv_code", b"spmv_off8", b"spmv_rowwin", b"spmv_rowmarch", b"spmv_zmarch")
    forms = (("zmarch", 1, 1, 1, 1, 1, 1), ("rowmarch", 1, 1, 1, 1, 1, 0), ("rowwin", 1, 1, 1, 1, 0, 0), ("rowpat", 1, 1, 1, 0, 0, 0),
             ("plain", 0, 0, 0, 0, 0, 0))
    N = n * n
    h = C.c_int64(0)
    kappa = np.ones((n + 2, n + 2))
    pkg.check(pkg.lib.ab_poisson2d_csr(n, kappa.ctypes.data, -1, C.byref(h)))
    A = pkg.SparseTensor.from_handle(h.value)
    gen = torch.Generator(device="cuda").manual_seed(233)
    x = torch.randn(N, dtype=torch.float64, device="cuda", generator=gen)
    ys = {}
    try:
        for name, *vals in forms:
            for k, v in zip(keys, vals):
                pkg.check(pkg.lib.ab_set_option(k, float(v)))
            y = torch.empty_like(x)
            pkg.check(pkg.lib.ab_spmv_solver_form(A.handle(), x.data_ptr(), y.data_ptr()))
            ys[name] = y
    finally:
        for k in keys:
            pkg.check(pkg.lib.ab_set_option(k, 1.0))
    for name, y in ys.items():
        assert torch.equal(y, ys["plain"]), (n, name)
    S = A.to_scipy()
    dinv = 1.0 / np.sqrt(S.diagonal())
    y0 = dinv * (S @ (dinv * x.cpu().numpy()))
    assert np.max(np.abs(ys["plain"].cpu().numpy() - y0)) <= 1e-13 * np.max(np.abs(y0))
    zi = C.c_double(0)
    pkg.check(pkg.lib.ab_csr_get_info(A.handle(), b"zmarch", C.byref(zi)))
    assert zi.value == 1.0
    # and 30 CG iterations through it agree with the plain kernel's iterates to rounding (fused dots differ in order)
    f = torch.rand(N, dtype=torch.float64, device="cuda", generator=gen)
    rr = C.c_double(0)
    us = {}
    try:
        for name, *vals in (forms[0], forms[-1]):
            for k, v in zip(keys, vals):
                pkg.check(pkg.lib.ab_set_option(k, float(v)))
            u = torch.empty_like(f)
            pkg.check(pkg.lib.ab_cg_fixed(A.handle(), f.data_ptr(), u.data_ptr(), 30, C.byref(rr)))
            us[name] = u
    finally:
        for k in keys:
            pkg.check(pkg.lib.ab_set_option(k, 1.0))
    assert float((us["zmarch"] - us["plain"]).abs().max() / us["plain"].abs().max()) <= 1e-11


def test_pattern_code_spmv_is_lossless(pkg):
    """The solver's compressed SpMV forms — row patterns (1 byte per row), pattern codes (1 byte per
    entry naming an {offset, value} pair), offset dictionary — must reproduce the plain int32 kernel
    BIT FOR BIT on y = (S A S) x, and must step aside by themselves when the values do not admit a
    256-entry dictionary (variable coefficients).  CG iterates of the different forms agree to rounding
    only (the fused inner products are summed in each kernel's own fixed tile order); every form is
    run-to-run deterministic."""
    import torch

    # name -> (spmv_rowpat, spmv_code, spmv_off8, spmv_rowwin, spmv_rowmarch, spmv_zmarch)
    forms = (("zmarch", 1.0, 1.0, 1.0, 1.0, 1.0, 1.0), ("rowmarch", 1.0, 1.0, 1.0, 1.0, 1.0, 0.0), ("rowwin", 1.0, 1.0, 1.0, 1.0, 0.0, 0.0),
             ("rowpat", 1.0, 1.0, 1.0, 0.0, 0.0, 0.0), ("code", 0.0, 1.0, 1.0, 0.0, 0.0, 0.0), ("off8", 0.0, 0.0, 1.0, 0.0, 0.0, 0.0),
             ("plain", 0.0, 0.0, 0.0, 0.0, 0.0, 0.0))
    keys = (b"spmv_rowpat", b"spmv_code", b"spmv_off8", b"spmv_rowwin", b"spmv_rowmarch", b"spmv_zmarch")

    def with_form(rowpat, code, off8, rowwin, rowmarch, zmarch, fn):
        for k, v in zip(keys, (rowpat, code, off8, rowwin, rowmarch, zmarch)):
            pkg.check(pkg.lib.ab_set_option(k, v))
        try:
            return fn()
        finally:
            for k in keys:
                pkg.check(pkg.lib.ab_set_option(k, 1.0))

    # 67: tiles straddle planes, odd sizes (x-window kernel); 64, 96, 128: n^2 is a multiple of the 1024-row tile ->
    # both marching kernels; 48, 100: n^2 is only a multiple of 16 -> the z-register kernel with the exact plane stride
    # (strips of the plane that do not coincide with tiles; the last strip is narrower)
    for n in (48, 67, 64, 96, 128, 100):
        N = n ** 3
        h = C.c_int64(0)
        pkg.check(pkg.lib.ab_poisson3d_csr(n, n, n, 0, N, C.byref(h)))
        A = pkg.SparseTensor.from_handle(h.value)
        gen = torch.Generator(device="cuda").manual_seed(233)
        f = torch.rand(N, dtype=torch.float64, device="cuda", generator=gen)
        x = torch.randn(N, dtype=torch.float64, device="cuda", generator=gen)
        rr = C.c_double(0)

        def spmv():
            y = torch.empty_like(x)
            pkg.check(pkg.lib.ab_spmv_solver_form(A.handle(), x.data_ptr(), y.data_ptr()))
            return y

        def cg():
            u = torch.empty_like(f)
            pkg.check(pkg.lib.ab_cg_fixed(A.handle(), f.data_ptr(), u.data_ptr(), 40, C.byref(rr)))
            return u

        ys = {name: with_form(a, b, c, d, e, g, spmv) for name, a, b, c, d, e, g in forms}
        if n in (48, 64, 96, 100, 128):
            zi = C.c_double(0)
            pkg.check(pkg.lib.ab_csr_get_info(A.handle(), b"zmarch", C.byref(zi)))
            assert zi.value == 1.0, "the z-register marching kernel should apply to this grid"
        pkg.check(pkg.lib.ab_set_option(b"spmv_rowwin_dom", 0.0))                # every row walks its pattern out of the stage
        try:
            ys["rowwin_generic"] = with_form(1.0, 1.0, 1.0, 1.0, 0.0, 0.0, spmv)
            ys["rowmarch_generic"] = with_form(1.0, 1.0, 1.0, 1.0, 1.0, 0.0, spmv)
            ys["zmarch_generic"] = with_form(1.0, 1.0, 1.0, 1.0, 1.0, 1.0, spmv)
        finally:
            pkg.check(pkg.lib.ab_set_option(b"spmv_rowwin_dom", 1.0))
        assert torch.equal(ys["rowwin_generic"], ys["plain"]) and torch.equal(ys["rowmarch_generic"], ys["plain"])
        assert torch.equal(ys["zmarch_generic"], ys["plain"])
        for name in ("zmarch", "rowmarch", "rowwin", "rowpat", "code", "off8"):
            assert torch.equal(ys[name], ys["plain"]), name                      # lossless: same bits
        S = A.to_scipy()
        dinv = 1.0 / np.sqrt(S.diagonal())
        y0 = dinv * (S @ (dinv * x.cpu().numpy()))
        assert np.max(np.abs(ys["plain"].cpu().numpy() - y0)) <= 1e-13 * np.max(np.abs(y0))
        us = {name: with_form(a, b, c, d, e, g, cg) for name, a, b, c, d, e, g in forms}
        for name in ("zmarch", "rowmarch", "rowwin", "rowpat", "code", "off8"):
            assert float((us[name] - us["plain"]).abs().max() / us["plain"].abs().max()) <= 1e-11, name
        assert torch.equal(with_form(1.0, 1.0, 1.0, 0.0, 0.0, 0.0, cg), us["rowpat"])     # deterministic run to run
        assert torch.equal(with_form(1.0, 1.0, 1.0, 1.0, 0.0, 0.0, cg), us["rowwin"])
        assert torch.equal(with_form(1.0, 1.0, 1.0, 1.0, 1.0, 0.0, cg), us["rowmarch"])
        assert torch.equal(with_form(1.0, 1.0, 1.0, 1.0, 1.0, 1.0, cg), us["zmarch"])
        pkg.check(pkg.lib.ab_set_option(b"spmv_rowwin", 0.0))               # the cg_defer_x checks below compare with us["rowpat"]
        # the x update riding in K3 (10 vector passes) is the same arithmetic, only moved
        pkg.check(pkg.lib.ab_set_option(b"cg_defer_x", 0.0))
        try:
            assert torch.equal(cg(), us["rowpat"])
            it, r2 = C.c_int(0), C.c_double(0)
            sols = []
            for d in (0.0, 1.0):               # and a converged solve stops with the last update applied
                pkg.check(pkg.lib.ab_set_option(b"cg_defer_x", d))
                u = torch.empty_like(f)
                pkg.check(pkg.lib.ab_solve(A.handle(), b"CG", f.data_ptr(), u.data_ptr(), 1e-9, 4000, C.byref(it), C.byref(r2)))
                sols.append((u, it.value))
            assert sols[0][1] == sols[1][1] and torch.equal(sols[0][0], sols[1][0])
        finally:
            pkg.check(pkg.lib.ab_set_option(b"cg_defer_x", 1.0))
            pkg.check(pkg.lib.ab_set_option(b"spmv_rowwin", 1.0))
        info = C.c_double(0)
        pkg.check(pkg.lib.ab_csr_get_info(A.handle(), b"rowwin", C.byref(info)))
        assert info.value == 1                 # the x-window form was built for the stencil
        pkg.check(pkg.lib.ab_csr_get_info(A.handle(), b"rowmarch", C.byref(info)))
        assert info.value == (1 if n in (64, 96, 128) else 0)      # chains need n^2 to be a multiple of the tile size
        pkg.check(pkg.lib.ab_csr_get_info(A.handle(), b"npat", C.byref(info)))
        assert info.value == 27                # interior + face / edge / corner variants of the 7-point stencil
        pkg.check(pkg.lib.ab_csr_get_info(A.handle(), b"ncode", C.byref(info)))
        assert info.value == 7
        A.close()
    # variable coefficients: > 256 distinct values -> the dictionary is refused, the solve is unchanged
    rng = np.random.default_rng(233)
    n = 64
    kext = np.exp(rng.uniform(-1, 1, (n + 2, n + 2)))
    h = C.c_int64(0)
    pkg.check(pkg.lib.ab_poisson2d_csr(n, kext.ctypes.data, -1, C.byref(h)))
    A = pkg.SparseTensor.from_handle(h.value)
    S = A.to_scipy()
    f = rng.uniform(size=n * n)
    u = pkg.backslash(A, f, "CG")
    assert np.linalg.norm(S @ u - f) / np.linalg.norm(f) <= TOL_SOLVE
    # new values on the same pattern invalidate the codes (values_version): constant -> scaled by 3
    h2 = C.c_int64(0)
    k_one, k_three = np.ones((n + 2, n + 2)), 3.0 * np.ones((n + 2, n + 2))      # kept alive: ctypes.data of a temporary dangles
    pkg.check(pkg.lib.ab_poisson2d_csr(n, k_one.ctypes.data, -1, C.byref(h2)))
    B = pkg.SparseTensor.from_handle(h2.value)
    u1 = pkg.backslash(B, f, "CG")
    pkg.check(pkg.lib.ab_poisson2d_update(h2.value, n, k_three.ctypes.data, -1))
    u3 = pkg.backslash(B, f, "CG")
    assert np.linalg.norm(3.0 * u3 - u1) / np.linalg.norm(u1) <= 1e-9


def test_window_spmv_matrix_classes(pkg):
    """The x-window SpMV (1024-row tiles staged by the TMA unit) against the plain int32 kernel, bit for bit, on
    matrix classes that exercise each of its paths: dominant patterns of 3 / 5 / 7 entries carried in the kernel
    parameters (1-D, 2-D, 3-D stencils), a 9-entry dominant pattern (every row walks its pattern out of the
    stage), odd vector lengths (the element the 16-byte bulk copy cannot take), windows clamped at both ends of
    the vector, a partial last tile, and a banded matrix with a few hundred row patterns."""
    import torch

    rng = np.random.default_rng(233)

    def stencil(shape, offsets_vals):
        N = int(np.prod(shape))
        idx = np.arange(N).reshape(shape)
        rows, cols, vals = [np.arange(N)], [np.arange(N)], [np.full(N, 4.0 + len(shape))]
        for d, v in offsets_vals:
            src = [slice(None)] * len(shape)
            dst = [slice(None)] * len(shape)
            ok = True
            for ax, dd in enumerate(d):
                if dd > 0: src[ax], dst[ax] = slice(0, shape[ax] - dd), slice(dd, None)
                elif dd < 0: src[ax], dst[ax] = slice(-dd, None), slice(0, shape[ax] + dd)
                ok = ok and abs(dd) < shape[ax]
            if not ok: continue
            r, c_ = idx[tuple(src)].ravel(), idx[tuple(dst)].ravel()
            rows.append(r); cols.append(c_); vals.append(np.full(r.size, v))
        return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N))

    cases = {
        "1d_3pt": stencil((5001,), [((1,), -1.0), ((-1,), -1.0)]),
        "2d_5pt": stencil((97, 131), [((0, 1), -1.0), ((0, -1), -1.0), ((1, 0), -1.0), ((-1, 0), -1.0)]),
        "3d_7pt_aniso": stencil((21, 33, 40), [((0, 0, 1), -1.0), ((0, 0, -1), -1.0), ((0, 1, 0), -0.5), ((0, -1, 0), -0.5),
                                                ((1, 0, 0), -0.25), ((-1, 0, 0), -0.25)]),
        "3d_7pt_march": stencil((8, 64, 64), [((0, 0, 1), -1.0), ((0, 0, -1), -1.0), ((0, 1, 0), -0.5), ((0, -1, 0), -0.5),
                                              ((1, 0, 0), -0.25), ((-1, 0, 0), -0.25)]),      # plane = 4 tiles: marching kernel
        "2d_9pt": stencil((64, 80), [((a, b), -0.5) for a in (-1, 0, 1) for b in (-1, 0, 1) if (a, b) != (0, 0)]),
        "tiny": stencil((7,), [((1,), -1.0), ((-1,), -1.0)]),
    }
    # banded SPD matrix with a handful of distinct values but hundreds of row patterns
    N = 6000
    B = sp.lil_matrix((N, N))
    for off in (1, 3, 17, 64):
        mask = rng.uniform(size=N - off) < 0.5
        v = np.where(mask, rng.choice([-0.25, -0.5], N - off), 0.0)
        B.setdiag(v, off); B.setdiag(v, -off)
    B.setdiag(np.full(N, 4.0))
    B = B.tocsr(); B.eliminate_zeros()
    cases["banded_many_patterns"] = B
    info = C.c_double(0)
    used = 0
    for name, S in cases.items():
        S = S.tocsr(); S.sort_indices()
        N = S.shape[0]
        A = pkg.SparseTensor(S)
        x = torch.randn(N, dtype=torch.float64, device="cuda")
        out = {}
        for use in (0.0, 1.0):
            pkg.check(pkg.lib.ab_set_option(b"spmv_rowwin", use))
            pkg.check(pkg.lib.ab_set_option(b"spmv_rowpat", use))
            pkg.check(pkg.lib.ab_set_option(b"spmv_code", use))
            pkg.check(pkg.lib.ab_set_option(b"spmv_off8", use))
            try:
                y = torch.empty_like(x)
                pkg.check(pkg.lib.ab_spmv_solver_form(A.handle(), x.data_ptr(), y.data_ptr()))
                out[use] = y
            finally:
                for k in (b"spmv_rowwin", b"spmv_rowpat", b"spmv_code", b"spmv_off8"):
                    pkg.check(pkg.lib.ab_set_option(k, 1.0))
        assert torch.equal(out[0.0], out[1.0]), name
        dinv = 1.0 / np.sqrt(S.diagonal())
        y0 = dinv * (S @ (dinv * x.cpu().numpy()))
        assert np.max(np.abs(out[1.0].cpu().numpy() - y0)) <= 1e-13 * np.max(np.abs(y0)), name
        pkg.check(pkg.lib.ab_csr_get_info(A.handle(), b"rowwin", C.byref(info)))
        used += int(info.value == 1)
        # a full solve through the window form agrees with scipy
        f = rng.uniform(size=N)
        u = pkg.backslash(A, f, "CG")
        assert np.linalg.norm(S @ u - f) / np.linalg.norm(f) <= TOL_SOLVE, name
        A.close()
    assert used >= 4, used


def test_dist_path_world1_equals_single_gpu(pkg):
    """ab_dist_* with one rank must reproduce the single-GPU solver bit for bit."""
    import torch

    n = 40
    N = n ** 3
    pkg.mpi_init(0, 1, torch.cuda.current_device())
    D = pkg.mpi_SparseTensor.poisson3d(n, n, n, 0, 1)
    x = torch.rand(N, dtype=torch.float64, device="cuda")
    assert torch.equal(D @ x, D.local @ x)
    b = D @ torch.ones(N, dtype=torch.float64, device="cuda")
    u, it, rr = D.solve(b, tol=1e-11)
    assert rr <= 1e-11 and float((u - 1).abs().max()) < 1e-8
    u1, _ = D.cg_fixed(b, 30)
    u2 = torch.empty_like(b)
    r = C.c_double(0)
    pkg.check(pkg.lib.ab_cg_fixed(D.local.handle(), b.data_ptr(), u2.data_ptr(), 30, C.byref(r)))
    assert torch.equal(u1, u2)
    # adjoint of the distributed solve == the single-GPU adjoint (per stored entry: the handle came from CSR)
    g = torch.rand(N, dtype=torch.float64, device="cuda")
    gv, gf = D.solve_grad(g, u, tol=1e-12)
    gv1 = torch.empty(D.local.query()[2], dtype=torch.float64, device="cuda")
    gf1 = torch.empty_like(g)
    pkg.check(pkg.lib.ab_solve_grad(D.local.handle(), b"CG", g.data_ptr(), u.data_ptr(), gv1.data_ptr(), gf1.data_ptr(),
                                    1e-12, 0, None, None))
    assert float((gf - gf1).abs().max() / gf1.abs().max()) <= 1e-9
    assert float((gv - gv1).abs().max() / gv1.abs().max()) <= 1e-9
    D.close()


def test_timers_and_launch_counter(pkg, orc):
    pkg.check(pkg.lib.ab_timer_reset())
    l0 = pkg.lib.ab_launch_count()
    ii, jj, vv, N = spd_case(orc, 30)
    A = pkg.SparseTensor(ii, jj, vv, N, N)
    pkg.backslash(A, np.ones(N), "CG")
    assert pkg.lib.ab_launch_count() > l0
    assert pkg.lib.ab_timer_get(0) > 0 and pkg.lib.ab_timer_get(2) > 0
