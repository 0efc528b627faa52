// cg_kernels.cuh — device-resident scalar block and the fused CG vector kernels, shared by the
// single-GPU solver (solver.cu) and the row-block distributed solver (dist.cu).
//
// `raw` mode (distributed): a kernel only deposits its local partial sums in SolveScal::red;
// the host-side driver all-reduces them with NCCL and runs cg_fin_kernel, which applies the same
// scalar logic the single-GPU kernels apply inside their last block.
#pragma once
#include "common.cuh"
#include <math.h>

namespace ab {

struct SolveScal {
    double rz[2];     // <r,z> ping-pong by iteration parity
    double dots[2];   // outputs of the SpMV-fused dots: [0] = <Ap,p> (CG) / <v,rhat>, <t,s> ; [1] = <y,y>
    double rr;        // <r,r>
    double bb;        // <b,b>
    double thr;       // tol^2 * bb ; negative disables the convergence test
    double alpha, omega, rho;
    double red[4];    // raw partial sums awaiting an all-reduce (distributed mode)
    int done;         // 1 converged / stopped
    int iters;        // iterations completed
    int breakdown;    // 1 = non-SPD / zero pivot / NaN
    int bad;          // K2 saw an unusable pAp (carried to cg_fin_kernel in raw mode)
};

constexpr int EW_THREADS = 256;

static inline unsigned ew_grid(size_t n) {
    size_t nb  = (n + EW_THREADS * 2 - 1) / (EW_THREADS * 2);
    size_t cap = (size_t)sm_count() * 8;
    if (cap > (size_t)MAX_REDUCE_GRID) cap = MAX_REDUCE_GRID;
    if (nb > cap) nb = cap;
    if (nb < 1) nb = 1;
    return (unsigned)nb;
}

__device__ __forceinline__ void cg_init_fin(SolveScal* s, double tol, double rz0, double bb) {
    s->rz[0] = rz0;
    s->rz[1] = 0.0;
    s->bb    = bb;
    s->rr    = bb;
    s->thr   = tol < 0 ? -1.0 : tol * tol * bb;
    s->iters = 0;
    s->bad   = 0;
    s->breakdown = (isfinite(rz0) && isfinite(bb)) ? 0 : 1;
    s->done  = (bb == 0.0 || s->breakdown) ? 1 : 0;   // b == 0 -> x = 0 is the solution
}
__device__ __forceinline__ void cg_update_fin(SolveScal* s, int parity, double rznew, double rr, bool bad) {
    s->rz[parity ^ 1] = rznew;
    s->rr = rr;
    s->iters += 1;
    if (bad || !isfinite(rznew) || !isfinite(rr)) { s->breakdown = 1; s->done = 1; }
    else if (s->thr >= 0.0 && rr <= s->thr) s->done = 1;
}

// r = b ; p = D^-1 b ; x = 0 ; rz = <r, D^-1 r> ; bb = <b,b>
static __global__ void __launch_bounds__(EW_THREADS)
cg_init_kernel(const double* __restrict__ b, const double* __restrict__ dinv, double* __restrict__ x,
               double* __restrict__ r, double* __restrict__ p, size_t n, double tol, SolveScal* s,
               double* partials, uint32_t* ticket, int raw) {
    __shared__ double sh[32];
    double a0 = 0.0, a1 = 0.0;
    for (size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * EW_THREADS) {
        double bi = b[i], z = dinv[i] * bi;
        x[i] = 0.0;
        r[i] = bi;
        p[i] = z;
        a0 = fma(bi, z, a0);
        a1 = fma(bi, bi, a1);
    }
    double mine[2];
    mine[0] = abd::block_sum<EW_THREADS>(a0, sh);
    mine[1] = abd::block_sum<EW_THREADS>(a1, sh);
    abd::grid_reduce<EW_THREADS, 2>(mine, partials, ticket, sh, [=](const double(&t)[2]) {
        if (raw) { s->red[0] = t[0]; s->red[1] = t[1]; s->done = 0; }
        else cg_init_fin(s, tol, t[0], t[1]);
    });
}

// K2: alpha = rz/pAp ; x += alpha p ; r -= alpha Ap ; rz' = <r, D^-1 r> ; rr = <r,r>
static __global__ void __launch_bounds__(EW_THREADS)
cg_update_kernel(const double* __restrict__ p, const double* __restrict__ Ap, const double* __restrict__ dinv,
                 double* __restrict__ x, double* __restrict__ r, size_t n, int parity, SolveScal* s,
                 double* partials, uint32_t* ticket, int raw) {
    __shared__ double sh[32];
    if (s->done) return;
    const double rz = s->rz[parity], pAp = s->dots[0];
    const bool bad = !(pAp > 0.0) || !isfinite(pAp) || !isfinite(rz);
    const double alpha = bad ? 0.0 : rz / pAp;
    double a0 = 0.0, a1 = 0.0;
    const size_t n2 = n >> 1;
    const double2* p2 = reinterpret_cast<const double2*>(p);
    const double2* Ap2 = reinterpret_cast<const double2*>(Ap);
    const double2* d2 = reinterpret_cast<const double2*>(dinv);
    double2* x2 = reinterpret_cast<double2*>(x);
    double2* r2 = reinterpret_cast<double2*>(r);
    for (size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < n2; i += (size_t)gridDim.x * EW_THREADS) {
        double2 pv = p2[i], av = Ap2[i], dv = d2[i], xv = x2[i], rv = r2[i];
        xv.x = fma(alpha, pv.x, xv.x);  xv.y = fma(alpha, pv.y, xv.y);
        rv.x = fma(-alpha, av.x, rv.x); rv.y = fma(-alpha, av.y, rv.y);
        x2[i] = xv;
        r2[i] = rv;
        a0 = fma(rv.x * dv.x, rv.x, a0); a0 = fma(rv.y * dv.y, rv.y, a0);
        a1 = fma(rv.x, rv.x, a1);        a1 = fma(rv.y, rv.y, a1);
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        size_t i = n - 1;
        double xv = fma(alpha, p[i], x[i]), rv = fma(-alpha, Ap[i], r[i]);
        x[i] = xv; r[i] = rv;
        a0 = fma(rv * dinv[i], rv, a0);
        a1 = fma(rv, rv, a1);
    }
    double mine[2];
    mine[0] = abd::block_sum<EW_THREADS>(a0, sh);
    mine[1] = abd::block_sum<EW_THREADS>(a1, sh);
    abd::grid_reduce<EW_THREADS, 2>(mine, partials, ticket, sh, [=](const double(&t)[2]) {
        s->alpha = alpha;
        if (raw) { s->red[0] = t[0]; s->red[1] = t[1]; s->bad = bad ? 1 : 0; }
        else cg_update_fin(s, parity, t[0], t[1], bad);
    });
}

// K3: beta = rz'/rz ; p = D^-1 r + beta p
static __global__ void __launch_bounds__(EW_THREADS)
cg_direction_kernel(const double* __restrict__ r, const double* __restrict__ dinv, double* __restrict__ p, size_t n,
                    int parity, const SolveScal* __restrict__ s) {
    if (s->done) return;
    const double beta = s->rz[parity ^ 1] / s->rz[parity];
    const size_t n2 = n >> 1;
    const double2* r2 = reinterpret_cast<const double2*>(r);
    const double2* d2 = reinterpret_cast<const double2*>(dinv);
    double2* p2 = reinterpret_cast<double2*>(p);
    for (size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < n2; i += (size_t)gridDim.x * EW_THREADS) {
        double2 rv = r2[i], dv = d2[i], pv = p2[i];
        pv.x = fma(beta, pv.x, rv.x * dv.x);
        pv.y = fma(beta, pv.y, rv.y * dv.y);
        p2[i] = pv;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        size_t i = n - 1;
        p[i] = fma(beta, p[i], r[i] * dinv[i]);
    }
}

// distributed mode: scalar logic after the all-reduce of SolveScal::red.  which: 0 init, 1 update
static __global__ void cg_fin_kernel(SolveScal* s, int which, int parity, double tol) {
    if (which == 0) cg_init_fin(s, tol, s->red[0], s->red[1]);
    else if (!s->done) cg_update_fin(s, parity, s->red[0], s->red[1], s->bad != 0);
}

// shared host helpers implemented in solver.cu
double* work_vec(Csr* c, int i);
int     read_scal(Csr* c, SolveScal* out);

}  // namespace ab
