"""GPU parity tests for the Newton-Raphson inner loop on the device:
ab_newton_step / ab_vec_* through the host mirror of src/optim.jl:462-647, against the reference's
known answers (test/optim.jl:235-249, test/extra.jl:28-49) and the oracle's restatement of the loop."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu


def test_vector_helpers(pkg):
    import torch

    rng = np.random.default_rng(233)
    for n in (1, 7, 1000, 100003):
        x, y = rng.standard_normal(n), rng.standard_normal(n)
        d = C.c_double(0)
        pkg.check(pkg.lib.ab_vec_dot(n, x.ctypes.data, y.ctypes.data, C.byref(d)))
        assert abs(d.value - float(x @ y)) <= 1e-12 * max(1.0, float(np.abs(x) @ np.abs(y)))
        out = np.empty(n)
        pkg.check(pkg.lib.ab_vec_axpby(n, 2.5, x.ctypes.data, -0.5, y.ctypes.data, out.ctypes.data))
        assert np.allclose(out, 2.5 * x - 0.5 * y, rtol=1e-14, atol=1e-14)      # fma(a, x, b*y): one rounding fewer
    xt = torch.tensor(x, device="cuda")
    d2 = C.c_double(0)
    pkg.check(pkg.lib.ab_vec_dot(n, xt.data_ptr(), xt.data_ptr(), C.byref(d2)))
    pkg.check(pkg.lib.ab_vec_dot(n, x.ctypes.data, x.ctypes.data, C.byref(d)))
    assert d.value == d2.value                       # fixed-order reduction: host or device operand, same bits


def test_newton_step_matches_formula(pkg, orc):
    rng = np.random.default_rng(233)
    n = 300
    A = sp.random(n, n, 0.02, random_state=3, format="coo") + sp.identity(n) * 5.0
    A = A.tocoo()
    J = pkg.SparseTensor(A.row + 1, A.col + 1, A.data, n, n)
    r, u = rng.standard_normal(n), rng.standard_normal(n)
    unew, delta = np.empty(n), np.empty(n)
    rn, li = C.c_double(0), C.c_int(0)
    pkg.check(pkg.lib.ab_newton_step(J.handle(), b"SparseLU", r.ctypes.data, u.ctypes.data, 0.75, unew.ctypes.data,
                                     delta.ctypes.data, C.byref(rn), 1e-13, 0, C.byref(li)))
    d0 = orc.sparse_solver(A.row + 1, A.col + 1, A.data, r)
    assert np.linalg.norm(delta - d0) <= 1e-10 * np.linalg.norm(d0)
    assert np.allclose(unew, u - 0.75 * delta, rtol=1e-14, atol=1e-14)
    assert abs(rn.value - np.linalg.norm(r)) <= 1e-13 * np.linalg.norm(r) and li.value > 0
    with pytest.raises(pkg.ADCMEError):
        R = pkg.SparseTensor(np.array([1]), np.array([2]), np.array([1.0]), 2, 3)
        pkg.check(pkg.lib.ab_newton_step(R.handle(), b"CG", r.ctypes.data, u.ctypes.data, 1.0, unew.ctypes.data, None,
                                         C.byref(rn), 0.0, 0, None))


def test_newton_raphson_cuberoot_golden_and_oracle_history(pkg, orc, golden):
    import torch

    g = golden["newton_cuberoot_adjoint"]                              # test/optim.jl:235-249
    theta = np.array(g["theta"], float)

    def f(th, x):                                                      # x^3 - θ, 3 spdiag(x^2)
        return x ** 3 - th, pkg.spdiag(3 * x ** 2)

    pkg.reset_default_options()
    nr = pkg.newton_raphson(f, np.ones(3), theta)
    x0, res0, U0, conv0, it0 = orc.newton_raphson(lambda th, x: (x ** 3 - th, sp.diags(3 * x ** 2)), np.ones(3), theta)
    assert np.allclose(nr.x, g["expect_x"], rtol=1e-12)
    assert nr.iter == it0 and nr.converged == conv0 and nr.u.shape == U0.shape
    assert np.allclose(nr.res, res0, rtol=1e-9, atol=1e-13) and np.allclose(nr.u, U0, rtol=1e-10)
    # device-resident iterate and the differentiable form
    th = torch.tensor(theta, device="cuda", requires_grad=True)
    x = pkg.newton_raphson_with_grad(f, torch.ones(3, dtype=torch.float64, device="cuda"), th)
    assert x.is_cuda and np.allclose(x.detach().cpu().numpy(), g["expect_x"], rtol=1e-12)
    x.sum().backward()
    assert np.allclose(th.grad.cpu().numpy(), g["expect_grad"], rtol=1e-10)
    # θ without gradient: plain forward ("θ is not a PyObject, no gradients is available")
    x2 = pkg.newton_raphson_with_grad(f, np.ones(3), theta)
    assert np.allclose(x2, g["expect_x"], rtol=1e-12)


def test_newton_register_example_golden(pkg, golden):
    import torch

    g = golden["register_newton_gradient"]                             # test/extra.jl:28-49
    xin = torch.full((5,), 8.0, dtype=torch.float64, device="cuda")
    th = torch.ones(5, dtype=torch.float64, device="cuda", requires_grad=True)

    def f(theta, y, x):
        return y ** 3 + 1.0 - theta - x, pkg.spdiag(3 * y ** 2)

    y = pkg.newton_raphson_with_grad(f, torch.ones(5, dtype=torch.float64, device="cuda"), th, xin)
    assert np.allclose(y.detach().cpu().numpy(), g["expect_y"], rtol=1e-12)
    y.sum().backward()
    assert np.allclose(th.grad.cpu().numpy(), g["expect_grad_each"], rtol=1e-10)


def test_newton_nonlinear_poisson_pattern_reuse_linesearch_and_lm(pkg, orc):
    """-Δu + u³ = f on a 24x24 grid: the Jacobian K + 3 diag(u²) keeps its index arrays, so the
    assembled handle is reused across iterations (values re-reduced, no re-sort)."""
    n = 24
    N = n * n
    ii, jj, vv = orc.facewise_triplets_2d(n, seed=233)
    rp, ci, va, _ = orc.coo_to_csr(ii, jj, vv, N, N, base=1)
    K = sp.csr_matrix((va, ci, rp), shape=(N, N))
    rng = np.random.default_rng(233)
    ustar = rng.uniform(0.5, 1.5, N)
    fr = K @ ustar + ustar ** 3
    dI = np.arange(1, N + 1, dtype=np.int64)
    II, JJ = np.concatenate([ii, dI]), np.concatenate([jj, dI])        # same index arrays every iteration

    def f(th, u):
        r = K @ u + u ** 3 - fr
        return 0.5 * float(r @ r), r, pkg.SparseTensor(II, JJ, np.concatenate([vv, 3 * u ** 2]), N, N)

    def f0(th, u):
        J = K + sp.diags(3 * u ** 2)
        return K @ u + u ** 3 - fr, J

    pkg.reset_default_options()
    pkg.options.newton_raphson.tol = 1e-9
    pkg.options.newton_raphson.rtol = 1e-9
    try:
        l0 = pkg.lib.ab_timer_get(2)
        nr = pkg.newton_raphson(f, np.ones(N), None)
        assembly_s = pkg.lib.ab_timer_get(2) - l0
        x0, res0, U0, conv0, it0 = orc.newton_raphson(f0, np.ones(N), None, tol=1e-9, rtol=1e-9)
        assert nr.converged and nr.iter == it0
        assert np.linalg.norm(nr.x - ustar) <= 1e-8 * np.linalg.norm(ustar)
        assert np.allclose(nr.res, res0, rtol=1e-6, atol=1e-11)
        assert assembly_s >= 0
        # backtracking line search and a Levenberg-Marquardt shift still converge to the same root
        pkg.options.newton_raphson.linesearch = True
        nr2 = pkg.newton_raphson(f, 3.0 * np.ones(N), None)
        assert np.linalg.norm(nr2.x - ustar) <= 1e-7 * np.linalg.norm(ustar)
        pkg.options.newton_raphson.linesearch = False
        pkg.options.newton_raphson.LM = 1e-3
        pkg.options.newton_raphson.max_iter = 200
        nr3 = pkg.newton_raphson(f, np.ones(N), None)
        x3, *_ = orc.newton_raphson(f0, np.ones(N), None, tol=1e-9, rtol=1e-9, LM=1e-3, max_iter=200)
        assert np.linalg.norm(nr3.x - x3) <= 1e-7 * np.linalg.norm(x3)
    finally:
        pkg.reset_default_options()
