// newton.cu — the device side of ADCME's newton_raphson loop.
//
// Reference: src/optim.jl:462-592 (newton_raphson), :364-426 (backtracking), :616-647
// (newton_raphson_with_grad).  The residual/Jacobian callback stays in the host language; what the
// reference runs per iteration through TensorFlow ops — `δ = J \ r`, `‖r‖`, `u - α δ`,
// `df0 = -Σ r·δ` — is what lives here, so the iterate, the residual and the Newton direction never
// leave the GPU between callbacks:
//
//   ab_newton_step   δ = J \ r (any `\` method), u_new = u - step·δ, ‖r‖₂  in one call
//   ab_vec_axpby     out = a·x + b·y          (line-search trial points u - α δ)
//   ab_vec_dot       <x, y>                    (df0, and ‖·‖₂² with y = x)
//
// Reductions are the fixed-order two-stage kind used everywhere else (no fp atomics), so a Newton
// history is bit-reproducible.
#include "common.cuh"
#include <math.h>

namespace ab {

constexpr int NV_THREADS = 256;

static unsigned nv_grid(size_t n) {
    size_t nb = (n + NV_THREADS * 2 - 1) / (NV_THREADS * 2), cap = (size_t)sm_count() * 8;
    if (cap > (size_t)MAX_REDUCE_GRID) cap = MAX_REDUCE_GRID;
    if (nb > cap) nb = cap;
    return (unsigned)(nb < 1 ? 1 : nb);
}

__global__ void __launch_bounds__(NV_THREADS)
vec_axpby_kernel(double a, const double* __restrict__ x, double b, const double* __restrict__ y, size_t n,
                 double* __restrict__ out) {
    for (size_t i = (size_t)blockIdx.x * NV_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * NV_THREADS)
        out[i] = fma(a, x[i], b * y[i]);
}

// res[0] = <x,y>
__global__ void __launch_bounds__(NV_THREADS)
vec_dot_kernel(const double* __restrict__ x, const double* __restrict__ y, size_t n, double* partials,
               uint32_t* ticket, double* res) {
    __shared__ double sh[32];
    double a = 0.0;
    for (size_t i = (size_t)blockIdx.x * NV_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * NV_THREADS)
        a = fma(x[i], y[i], a);
    double mine[1];
    mine[0] = abd::block_sum<NV_THREADS>(a, sh);
    abd::grid_reduce<NV_THREADS, 1>(mine, partials, ticket, sh, [=](const double(&t)[1]) { res[0] = t[0]; });
}

// u_new = u - step*delta ; res[0] = <r,r>
__global__ void __launch_bounds__(NV_THREADS)
newton_update_kernel(const double* __restrict__ u, const double* __restrict__ delta, const double* __restrict__ r,
                     double step, size_t n, double* __restrict__ unew, double* partials, uint32_t* ticket, double* res) {
    __shared__ double sh[32];
    double a = 0.0;
    for (size_t i = (size_t)blockIdx.x * NV_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * NV_THREADS) {
        unew[i] = fma(-step, delta[i], u[i]);
        a = fma(r[i], r[i], a);
    }
    double mine[1];
    mine[0] = abd::block_sum<NV_THREADS>(a, sh);
    abd::grid_reduce<NV_THREADS, 1>(mine, partials, ticket, sh, [=](const double(&t)[1]) { res[0] = t[0]; });
}

struct ReduceScratch {
    DevBuf<double> partials, res;
    DevBuf<uint32_t> ticket;
    int init() {
        AB_TRY(partials.alloc(MAX_REDUCE_GRID)); AB_TRY(res.alloc(1)); AB_TRY(ticket.alloc(TICKET_WORDS));
        AB_CUDA(cudaMemsetAsync(ticket.p, 0, TICKET_WORDS * sizeof(uint32_t), stream()));
        return AB_OK;
    }
    int fetch(double* out) {
        AB_CUDA(cudaMemcpyAsync(out, res.p, sizeof(double), cudaMemcpyDeviceToHost, stream()));
        AB_CUDA(cudaStreamSynchronize(stream()));
        return AB_OK;
    }
};

}  // namespace ab

using namespace ab;

extern "C" {

int ab_vec_axpby(int64_t n, double a, const double* x_, double b, const double* y_, double* out_) {
    AB_TRY(ensure_device());
    if (n < 0) return fail(AB_EINVAL, "negative size");
    if (n == 0) return AB_OK;
    if (!x_ || !y_ || !out_) return fail(AB_EINVAL, "null vector");
    std::lock_guard<std::mutex> lk(big_lock());
    In<double> x, y; Out<double> out;
    AB_TRY(x.bind(x_, n)); AB_TRY(y.bind(y_, n)); AB_TRY(out.bind(out_, n));
    vec_axpby_kernel<<<nv_grid((size_t)n), NV_THREADS, 0, stream()>>>(a, x.d, b, y.d, (size_t)n, out.d);
    AB_LAUNCH_CHECK();
    AB_TRY(out.commit());
    if (x.tmp.p || y.tmp.p) AB_CUDA(cudaStreamSynchronize(stream()));
    return AB_OK;
}

int ab_vec_dot(int64_t n, const double* x_, const double* y_, double* result) {
    AB_TRY(ensure_device());
    if (n < 0 || !result) return fail(AB_EINVAL, "bad argument");
    *result = 0.0;
    if (n == 0) return AB_OK;
    if (!x_ || !y_) return fail(AB_EINVAL, "null vector");
    std::lock_guard<std::mutex> lk(big_lock());
    In<double> x, y;
    AB_TRY(x.bind(x_, n)); AB_TRY(y.bind(y_, n));
    ReduceScratch rs;
    AB_TRY(rs.init());
    vec_dot_kernel<<<nv_grid((size_t)n), NV_THREADS, 0, stream()>>>(x.d, y.d, (size_t)n, rs.partials.p, rs.ticket.p, rs.res.p);
    AB_LAUNCH_CHECK();
    return rs.fetch(result);
}

// One Newton-Raphson update (src/optim.jl:509-539): delta = J \ r ; u_new = u - step*delta ; rnorm = ||r||_2.
// h is the assembled Jacobian (LM shifts are added by the caller as J + mu*I, optim.jl:520-524).
// delta (optional output) is what the line search needs; u_new may alias u.
int ab_newton_step(int64_t h, const char* method, const double* r_, const double* u_, double step,
                   double* unew_, double* delta_, double* rnorm, double tol, int maxit, int* lin_iters) {
    AB_TRY(ensure_device());
    int64_t m = 0, n = 0;
    AB_TRY(ab_csr_query(h, &m, &n, nullptr, nullptr));
    if (m != n) return fail(AB_EINVAL, "newton_step: the Jacobian must be square (%lld x %lld)", (long long)m, (long long)n);
    if (m > 0 && (!r_ || !u_ || !unew_)) return fail(AB_EINVAL, "null vector");
    if (m == 0) { if (rnorm) *rnorm = 0.0; return AB_OK; }
    In<double> r, u; Out<double> unew, delta;
    DevBuf<double> dtmp;
    {
        std::lock_guard<std::mutex> lk(big_lock());
        AB_TRY(r.bind(r_, m)); AB_TRY(u.bind(u_, m)); AB_TRY(unew.bind(unew_, m));
        if (delta_) AB_TRY(delta.bind(delta_, m));
        else { AB_TRY(dtmp.alloc(m)); delta.d = dtmp.p; }
    }
    AB_TRY(ab_solve(h, method, r.d, delta.d, tol, maxit, lin_iters, nullptr));   // device pointers: no staging inside
    std::lock_guard<std::mutex> lk(big_lock());
    ReduceScratch rs;
    AB_TRY(rs.init());
    newton_update_kernel<<<nv_grid((size_t)m), NV_THREADS, 0, stream()>>>(u.d, delta.d, r.d, step, (size_t)m, unew.d, rs.partials.p,
                                                                         rs.ticket.p, rs.res.p);
    AB_LAUNCH_CHECK();
    double rr = 0.0;
    AB_TRY(rs.fetch(&rr));
    if (rnorm) *rnorm = sqrt(rr);
    AB_TRY(unew.commit());
    if (delta_) AB_TRY(delta.commit());
    return AB_OK;
}

}  // extern "C"
