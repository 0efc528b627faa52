"""Host-side mirror of ADCME's Newton-Raphson driver (/root/reference/src/optim.jl:351-647) over the
C-ABI of csrc/newton.cu.

The residual/Jacobian callback `func(θ, u, *args) -> (r, J)` or `(val, r, J)` stays in the host
language, exactly as in the reference (there it is Julia code building TF ops); everything the
reference then does with TF ops per iteration — `δ = J \\ r`, `‖r‖`, `u - α δ`, the line-search
reductions — is one call into libadcme_b200.so on device-resident vectors.  When successive
Jacobians share their index arrays the assembled CSR handle is reused and only its values are
re-reduced (ab_csr_update_values) — the `factorize`-style reuse the reference documents for
while-loops (docs/src/factorization.md:23-73).

    NRResult                      optim.jl:351-357
    backtracking                  optim.jl:364-426
    newton_raphson                optim.jl:462-592
    newton_raphson_with_grad      optim.jl:616-647
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from . import _lib, sparse as _sp, structure as _st
from ._lib import as_array, check, empty_like_kind, lib, ptr
from .options import options


def _is_torch(x) -> bool:
    return _lib._is_torch(x)


@dataclass
class NRResult:
    """optim.jl:351-357"""
    x: object            # final solution
    res: np.ndarray      # residual history
    u: object            # solution history, one COLUMN per iterate (the reference returns u_his')
    converged: bool
    iter: int


def _vec(x):
    """float64 vector living where it already lives (numpy / torch cpu / torch cuda)."""
    if _is_torch(x):
        import torch

        return x.detach().to(torch.float64).contiguous()
    return np.ascontiguousarray(x, dtype=np.float64)


def _dot(x, y) -> float:
    kx, px = as_array(x, np.float64)
    ky, py = as_array(y, np.float64)
    out = C.c_double(0)
    check(lib.ab_vec_dot(_sp._numel(kx), px, py, C.byref(out)))
    return out.value


def _norm(x) -> float:
    return math.sqrt(_dot(x, x))


def _axpby(a, x, b, y):
    kx, px = as_array(x, np.float64)
    ky, py = as_array(y, np.float64)
    out = empty_like_kind(kx, (_sp._numel(kx),), np.float64)
    check(lib.ab_vec_axpby(_sp._numel(kx), float(a), px, float(b), py, ptr(out)))
    return out


def _same_indices(a, b) -> bool:
    if a is b:
        return True
    if a is None or b is None or tuple(a.shape) != tuple(b.shape) or _is_torch(a) != _is_torch(b):
        return False
    if _is_torch(a):
        import torch

        return a.device == b.device and bool(torch.equal(a, b))
    return bool(np.array_equal(a, b))


class _JacobianCache:
    """Keeps the assembled CSR of the previous Jacobian; a new Jacobian on the same index arrays
    only refreshes the values (no re-sort)."""

    def __init__(self):
        self.J = None

    def adopt(self, J):
        if not isinstance(J, _sp.SparseTensor):
            J = _sp.SparseTensor(J)
        P = self.J
        if (P is not None and P is not J and not J._pattern_only and not P._pattern_only and P.shape == J.shape
                and _same_indices(P.ii, J.ii) and _same_indices(P.jj, J.jj)):
            P._refresh_values(J.vv)
            P.vv = J.vv
            return P
        self.J = J
        return J


def _call(func, theta, u, args, kwargs):
    """-> (val or None, r, J)   (optim.jl:512-518)"""
    out = func(theta, u, *args, **kwargs)
    if len(out) == 2:
        return None, out[0], out[1]
    return out[0], out[1], out[2]


def _newton_direction(cache, J, r, u, step, method):
    """δ = (J + μI) \\ r ; u_new = u - step δ ; ‖r‖   (optim.jl:519-539) — one ab_newton_step call."""
    mu = options.newton_raphson.LM
    if mu > 0.0:                                   # Levenberg-Marquardt: J + μ*spdiag(n), optim.jl:520-524
        J = _st.add(J, _st.spdiag(J.m), 1.0, float(mu))
    J = cache.adopt(J)
    kr, pr = as_array(r, np.float64)
    ku, pu = as_array(u, np.float64)
    unew = empty_like_kind(ku, (J.m,), np.float64)
    delta = empty_like_kind(ku, (J.m,), np.float64)
    rnorm = C.c_double(0)
    check(lib.ab_newton_step(J.handle(), method.encode(), pr, pu, float(step), ptr(unew), ptr(delta), C.byref(rnorm),
                             options.sparse.tol, options.sparse.maxiter, None))
    return unew, delta, rnorm.value


def backtracking(compute, u, f0, r0, delta0):
    """Backtracking line search (optim.jl:364-426).  `compute(x)` returns the merit value at x;
    f0, r0, δ0 are the merit value, residual and Newton direction at u.  Returns the step size α."""
    ls = options.newton_raphson.linesearch_options
    if f0 is None:
        raise ValueError("linesearch needs a callback returning (val, r, J)")        # @assert !isnothing(f0)
    assert ls.rho_lo < ls.rho_hi and ls.iterations > 0
    f0 = float(f0)
    df0 = -_dot(r0, delta0)                                                          # optim.jl:366
    alphas = {1: ls.alpha_initial, 2: ls.alpha_initial}
    fs = {1: 0.0, 2: f0}
    i = 2
    while fs[i] > f0 + ls.c1 * alphas[i] * df0 and i <= ls.iterations:               # optim.jl:378-382
        a1, a2, f = alphas[i - 1], alphas[i], fs[i]
        with np.errstate(all="ignore"):
            d = np.float64(1.0) / np.float64(a1 ** 2 * a2 ** 2 * (a2 - a1))          # inf when α₁ == α₂, as in the reference
            a = (a1 ** 2 * (f - f0 - df0 * a2) - a2 ** 2 * (df0 - f0 - df0 * a1)) * d
            b = (-a1 ** 3 * (f - f0 - df0 * a2) + a2 ** 3 * (df0 - f0 - df0 * a1)) * d
            if abs(a) < 1e-10:
                atmp = df0 / (2 * b)
            else:
                atmp = (-b + math.sqrt(max(b * b - 3 * a * df0, 0.0))) / (3 * a)
        if math.isnan(atmp):
            a2 = a2 * ls.rho_hi
        else:
            atmp = min(atmp, a2 * ls.rho_hi)
            a2 = max(atmp, a2 * ls.rho_lo)
        fnew = compute(_axpby(1.0, u, -a2, delta0))                                  # u - α₂ δ₀
        fs[i + 1], alphas[i + 1] = float(fnew), a2
        i += 1
    return alphas[i]


def newton_raphson(func, u0, theta=None, *args, method: str = "auto", **kwargs) -> NRResult:
    """newton_raphson(func, u0, θ, args...) (optim.jl:462-592).  Termination follows the reference
    loop literally: iteration i continues while ‖r(u_{i-1})‖ >= tol and ‖r(u_{i-2})‖ >= rtol and
    i <= max_iter (the reference's `rel_tol` is the previous residual norm, optim.jl:487-496)."""
    nr = options.newton_raphson
    if len(tuple(u0.shape)) != 1:
        raise ValueError("ADCME: Initial guess must be a vector")
    u = _vec(u0)
    method = _sp._resolve_method(method)
    cache = _JacobianCache()
    _, r0, _ = _call(func, theta, u, args, kwargs)
    tol0 = _norm(_vec(r0))
    ta_r = {1: tol0}
    ta_u = {1: u}
    i = 2
    while True:
        if i == 2:                                                                   # optim.jl:485-497
            cont = i < nr.max_iter + 1
        else:
            cont = ta_r[i - 1] >= nr.tol and ta_r[i - 2] >= nr.rtol and i <= nr.max_iter
        if not cont:
            break
        u_ = ta_u[i - 1]
        val, r_, J = _call(func, theta, u_, args, kwargs)
        r_ = _vec(r_)
        if nr.linesearch:                                                            # optim.jl:529-539
            _, d0, rnorm = _newton_direction(cache, J, r_, u_, 0.0, method)          # one solve per iteration
            step = backtracking(lambda x: _call(func, theta, x, args, kwargs)[0], u_, val, r_, d0)
            new_u = _axpby(1.0, u_, -step, d0)
            if nr.verbose:
                print("Current Step Size = ", step)
        else:
            new_u, _, rnorm = _newton_direction(cache, J, r_, u_, 1.0, method)
        ta_r[i] = rnorm
        ta_u[i] = new_u
        if nr.verbose:
            print(f"Newton iteration {i - 2}: r (abs) = {ta_r[i]} r (rel) = {ta_r[i - 1]}")
        i += 1
    i_ = i
    if tol0 < nr.tol:                                                                # optim.jl:560-584
        sol, res, his, it = u, np.array([tol0]), [u], 1
    else:
        sol = ta_u[i_ - 1]
        res = np.array([ta_r[k] for k in range(2, i_)])                              # tf.slice(r_out, [1], [i_-2])
        his = [ta_u[k] for k in range(1, i_ - 1)]                                    # tf.slice(u_out, [0,0], [i_-2, n])
        it = i_ - 2
    converged = i_ < nr.max_iter
    if _is_torch(sol):
        import torch

        U = torch.stack(his, dim=1) if his else torch.empty((sol.numel(), 0), dtype=sol.dtype, device=sol.device)
    else:
        U = np.stack(his, axis=1) if his else np.empty((len(sol), 0))
    return NRResult(sol, res, U, converged, it)


def newton_raphson_with_grad(func, u0, theta=None, *args, method: str = "auto", **kwargs):
    """Differentiable Newton-Raphson (optim.jl:616-647): forward = newton_raphson(...).x ; backward
    g = A'⁻¹ dy (A = Jacobian at the solution, not differentiated), s = Σ r·g, grads = -∂s/∂θ, -∂s/∂args."""
    if not (_is_torch(theta) and theta.requires_grad) and not any(_is_torch(a) and a.requires_grad for a in args):
        return newton_raphson(func, u0, theta, *args, method=method, **kwargs).x     # "no gradients is available"
    import torch

    meth = _sp._resolve_method(method)

    class NRFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, th, *xs):
            with torch.no_grad():
                x = newton_raphson(func, u0, th, *xs, method=method, **kwargs).x
            if not _is_torch(x):
                x = torch.as_tensor(x, device=th.device)
            ctx.save_for_backward(x, th, *xs)
            return x

        @staticmethod
        def backward(ctx, dy):
            x, th, *xs = ctx.saved_tensors
            with torch.enable_grad():
                th2 = th.detach().clone().requires_grad_(th.requires_grad)
                xs2 = [z.detach().clone().requires_grad_(z.requires_grad) for z in xs]
                out = func(th2, x.detach(), *xs2, **kwargs)
                r, A = (out[0], out[1]) if len(out) == 2 else (out[1], out[2])
                if not isinstance(A, _sp.SparseTensor):
                    A = _sp.SparseTensor(A)
                # g = independent(A' \\ dy): x_adj of ab_solve_grad IS A^{-T} dy
                dyc = dy.detach().contiguous()
                g = torch.empty_like(dyc)
                check(lib.ab_solve_grad(A.handle(), meth.encode(), ptr(dyc), ptr(dyc), None, ptr(g),
                                        options.sparse.tol, options.sparse.maxiter, None, None))
                s = (r * g).sum()
                wanted = [z for z in [th2] + xs2 if z.requires_grad]
                grads = torch.autograd.grad(s, wanted, allow_unused=True)
            it = iter(grads)
            outg = []
            for z in [th2] + xs2:
                gz = next(it) if z.requires_grad else None
                outg.append(None if gz is None else -gz)
            return tuple(outg)

    return NRFn.apply(theta, *args)
