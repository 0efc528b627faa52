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
    int half;         // BiCGSTAB: half-step flag / carried `bad` (raw mode)
};

constexpr int EW_THREADS = 256;

static inline unsigned ew_grid(size_t n) {
    size_t nb  = (n + EW_THREADS * 2 - 1) / (EW_THREADS * 2);
    // measured (profiles/slab_floor_r02l.log, 512^3): 4 CTAs of 256 threads per SM -> update 0.5015 / direction 0.8384 ms;
    // 8 -> 0.5128 / 0.8736; 3, 5, 6 are worse again (the grid-stride distance interacts with the DRAM channel interleave)
    int per_sm = (int)opt("ew_ctas_per_sm", 4.0);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    size_t cap = (size_t)sm_count() * per_sm;
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

// Residual replacement (van der Vorst & Ye): every `replace_every` iterations of a LONG solve the
// recurrence residual is overwritten by the true one, r = b - A x, so it cannot drift away from
// it on ill-conditioned systems (kappa ~ 1e7 at 4096^2: drift 1e-10 -> 1.6e-8 without this).
// Ax holds A x (scaled system when sb != nullptr: r~ = S b - A~ x~).  Overrides rz', rr of K2.
static __global__ void __launch_bounds__(EW_THREADS)
cg_replace_kernel(const double* __restrict__ b, const double* __restrict__ sb, const double* __restrict__ Ax,
                  const double* __restrict__ dz, double* __restrict__ r, size_t n, int parity, SolveScal* s,
                  double* partials, uint32_t* ticket) {
    __shared__ double sh[32];
    if (s->done) return;
    double a0 = 0.0, a1 = 0.0;
    for (size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * EW_THREADS) {
        const double ri = (sb ? b[i] * sb[i] : b[i]) - Ax[i];
        r[i] = ri;
        a0 = fma(dz ? ri * dz[i] : ri, ri, a0);
        a1 = fma(ri, ri, a1);
    }
    double mine[2];
    mine[0] = abd::block_sum<EW_THREADS>(a0, sh);
    mine[1] = abd::block_sum<EW_THREADS>(a1, sh);
    abd::grid_reduce<EW_THREADS, 2>(mine, partials, ticket, sh, [=](const double(&t)[2]) {
        s->rz[parity ^ 1] = t[0];
        s->rr = t[1];
        if (!isfinite(t[0]) || !isfinite(t[1])) { s->breakdown = 1; s->done = 1; }
        else if (s->thr >= 0.0 && t[1] <= s->thr) s->done = 1;
    });
}

// distributed mode: scalar logic after the all-reduce of SolveScal::red.  which: 0 init, 1 update
static __global__ void cg_fin_kernel(SolveScal* s, int which, int parity, double tol);

// ---------------------------------------------------------------------------------------------
// Symmetrically scaled form of Jacobi-PCG.  With S = D^-1/2, PCG on (A, M = D) is CG on
// A~ = S A S, x~ = S^-1 x, b~ = S b (same iterates up to rounding).  The preconditioner
// disappears from the vector kernels: K2 streams 6 vectors instead of 7, K3 3 instead of 4, and
// one inner product (<r~,r~>) serves both the recurrence and the stopping test:
//     12*nnz + 4*(N+1) + 11*8*N bytes per iteration instead of ... + 13*8*N;
// with the x update deferred into K3 (defer_x) K2 streams 3 and K3 5: 10*8*N.
// The stopping test is on ||r~||/||b~|| (= ||r||/||b|| for a constant diagonal); the caller's
// guarantee on the TRUE residual comes from the refinement step in solve_dev.
// ---------------------------------------------------------------------------------------------
// r = S b ; p = r ; x = 0 ; rz = bb = <r,r>
static __global__ void __launch_bounds__(EW_THREADS)
cgs_init_kernel(const double* __restrict__ b, const double* __restrict__ sinv, double* __restrict__ x,
                double* __restrict__ r, double* __restrict__ p, size_t n, double tol, SolveScal* s,
                double* partials, uint32_t* ticket, int raw) {
    __shared__ double sh[32];
    double a0 = 0.0;
    for (size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * EW_THREADS) {
        const double ri = b[i] * sinv[i];
        x[i] = 0.0;
        r[i] = ri;
        p[i] = ri;
        a0 = fma(ri, ri, a0);
    }
    double mine[2];
    mine[0] = abd::block_sum<EW_THREADS>(a0, sh);
    mine[1] = mine[0];
    abd::grid_reduce<EW_THREADS, 2>(mine, partials, ticket, sh, [=](const double(&t)[2]) {
        if (raw) { s->red[0] = t[0]; s->red[1] = t[1]; s->done = 0; }
        else cg_init_fin(s, tol, t[0], t[1]);
    });
}

// K2: alpha = rz/pAp ; r -= alpha Ap ; rz' = rr = <r,r> ; and, unless defer_x, x += alpha p.
// defer_x != 0 moves the x update into K3, which streams p anyway: K2 then touches 3 vectors and K3 5,
// i.e. 10 instead of 11 vector passes per iteration (12*nnz-equivalent matrix bytes + 10*8*N).
static __global__ void __launch_bounds__(EW_THREADS)
cgs_update_kernel(const double* __restrict__ p, const double* __restrict__ Ap, double* __restrict__ x,
                  double* __restrict__ r, size_t n, int parity, SolveScal* s, double* partials, uint32_t* ticket,
                  int raw, int defer_x, ArExch ex) {
    __shared__ double sh[32];
    abd::pdl_launch_dependents();
    abd::pdl_wait();
    if (s->done) return;
    const double rz = s->rz[parity], pAp = s->dots[0];
    const bool bad = !(pAp > 0.0) || !isfinite(pAp) || !isfinite(rz);
    const double alpha = bad ? 0.0 : rz / pAp;
    double a0 = 0.0;
    const size_t n2 = n >> 1;
    const double2* p2 = reinterpret_cast<const double2*>(p);
    const double2* Ap2 = reinterpret_cast<const double2*>(Ap);
    double2* x2 = reinterpret_cast<double2*>(x);
    double2* r2 = reinterpret_cast<double2*>(r);
    if (defer_x) {
        // two grid-stride steps per trip: four 16-byte loads in flight per thread (at 4 CTAs per SM the single-step loop
        // keeps 32 KB per SM in flight, short of what HBM latency needs); the sum order per thread is unchanged
        // — measured (profiles/slab_floor_r02s.log): 73.8 -> 70.0 us on a 16.7 M-row slab, but 501.5 -> 511.7 us at 134 M rows
        // (the doubled stride meets the DRAM channel interleave badly), so only vectors up to 48 M rows take it
        const size_t step = (size_t)gridDim.x * EW_THREADS;
        size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x;
        const size_t n2u = n <= ((size_t)48 << 20) ? n2 : 0;
        for (; i + step < n2u; i += 2 * step) {
            double2 av = Ap2[i], rv = r2[i], av1 = Ap2[i + step], rv1 = r2[i + step];
            rv.x = fma(-alpha, av.x, rv.x); rv.y = fma(-alpha, av.y, rv.y);
            rv1.x = fma(-alpha, av1.x, rv1.x); rv1.y = fma(-alpha, av1.y, rv1.y);
            r2[i] = rv;
            r2[i + step] = rv1;
            a0 = fma(rv.x, rv.x, a0); a0 = fma(rv.y, rv.y, a0);
            a0 = fma(rv1.x, rv1.x, a0); a0 = fma(rv1.y, rv1.y, a0);
        }
        for (; i < n2; i += step) {
            double2 av = Ap2[i], rv = r2[i];
            rv.x = fma(-alpha, av.x, rv.x); rv.y = fma(-alpha, av.y, rv.y);
            r2[i] = rv;
            a0 = fma(rv.x, rv.x, a0); a0 = fma(rv.y, rv.y, a0);
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < n2; i += (size_t)gridDim.x * EW_THREADS) {
            double2 pv = p2[i], av = Ap2[i], xv = x2[i], rv = r2[i];
            xv.x = fma(alpha, pv.x, xv.x);  xv.y = fma(alpha, pv.y, xv.y);
            rv.x = fma(-alpha, av.x, rv.x); rv.y = fma(-alpha, av.y, rv.y);
            x2[i] = xv;
            r2[i] = rv;
            a0 = fma(rv.x, rv.x, a0); a0 = fma(rv.y, rv.y, a0);
        }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        size_t i = n - 1;
        double rv = fma(-alpha, Ap[i], r[i]);
        if (!defer_x) x[i] = fma(alpha, p[i], x[i]);
        r[i] = rv;
        a0 = fma(rv, rv, a0);
    }
    double mine[2];
    mine[0] = abd::block_sum<EW_THREADS>(a0, sh);
    mine[1] = mine[0];
    abd::grid_reduce<EW_THREADS, 2>(mine, partials, ticket, sh, [=](const double(&t)[2]) {
        s->alpha = alpha;
        if (raw && ex.blk) {
            // the box-wide <r,r> right here (one thread of the last CTA): no all-reduce launch between K2 and K3
            double g0 = t[0], g1 = t[1];
            abd::ll_allreduce2(ex.blk, ex.peer_blk, ex.world, ex.me, ex.seq, g0, g1);
            cg_update_fin(s, parity, g0, g1, bad);
        } else if (raw) { s->red[0] = t[0]; s->red[1] = t[1]; s->bad = bad ? 1 : 0; }
        else cg_update_fin(s, parity, t[0], t[1], bad);
    });
}

// K3: beta = rz'/rz ; p = r + beta p ; with defer_x also x += alpha p_old (the update K2 left out).
// `it` is the 0-based iteration this launch belongs to: s->iters == it + 1 says K2 of this iteration
// really ran (it did not early-exit on `done`), so its x update is still owed even when that same K2
// detected convergence; launches of later iterations see iters unchanged and do nothing.
static __global__ void __launch_bounds__(EW_THREADS)
cgs_direction_kernel(const double* __restrict__ r, double* __restrict__ p, double* __restrict__ x, size_t n,
                     int parity, int it, int defer_x, const SolveScal* __restrict__ s) {
    abd::pdl_launch_dependents();
    abd::pdl_wait();
    const bool upd_x = defer_x && (s->iters == it + 1);
    const bool upd_p = !s->done;
    if (!upd_x && !upd_p) return;
    const double alpha = s->alpha;
    const double beta = upd_p ? s->rz[parity ^ 1] / s->rz[parity] : 0.0;
    const size_t n2 = n >> 1;
    const double2* r2 = reinterpret_cast<const double2*>(r);
    double2* p2 = reinterpret_cast<double2*>(p);
    double2* x2 = reinterpret_cast<double2*>(x);
    if (upd_x && upd_p) {
        for (size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < n2; i += (size_t)gridDim.x * EW_THREADS) {
            double2 rv = r2[i], pv = p2[i], xv = x2[i];
            xv.x = fma(alpha, pv.x, xv.x); xv.y = fma(alpha, pv.y, xv.y);
            pv.x = fma(beta, pv.x, rv.x);  pv.y = fma(beta, pv.y, rv.y);
            x2[i] = xv;
            p2[i] = pv;
        }
    } else if (upd_p) {
        for (size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < n2; i += (size_t)gridDim.x * EW_THREADS) {
            double2 rv = r2[i], pv = p2[i];
            pv.x = fma(beta, pv.x, rv.x);
            pv.y = fma(beta, pv.y, rv.y);
            p2[i] = pv;
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < n2; i += (size_t)gridDim.x * EW_THREADS) {
            double2 pv = p2[i], xv = x2[i];
            xv.x = fma(alpha, pv.x, xv.x); xv.y = fma(alpha, pv.y, xv.y);
            x2[i] = xv;
        }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        size_t i = n - 1;
        const double pv = p[i];
        if (upd_x) x[i] = fma(alpha, pv, x[i]);
        if (upd_p) p[i] = fma(beta, pv, r[i]);
    }
}

// x = S x~
static __global__ void scale_vec_kernel(double* __restrict__ x, const double* __restrict__ sinv, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] *= sinv[i];
}

// sinv[r] = 1/sqrt(diag); *bad set when a diagonal entry is missing or not positive (not SPD)
static __global__ void scale_diag_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                                         const double* __restrict__ values, int64_t m, double* __restrict__ sinv,
                                         int* __restrict__ bad) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    double d = 0.0;
    for (int j = rowptr[r]; j < rowptr[r + 1]; ++j)
        if (colind[j] == r) d += values[j];
    if (!(d > 0.0) || !isfinite(d)) { atomicExch(bad, 1); sinv[r] = 1.0; }
    else sinv[r] = 1.0 / sqrt(d);
}
// sv[e] = v[e] * sinv[row] * sinv_ext[col]   (sinv_ext covers halo columns in the distributed case)
static __global__ void scale_values_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                                           const double* __restrict__ values, int64_t m,
                                           const double* __restrict__ sinv_ext, double* __restrict__ sv) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const double sr = sinv_ext[r];
    for (int j = rowptr[r]; j < rowptr[r + 1]; ++j) sv[j] = values[j] * sr * sinv_ext[colind[j]];
}

// ---------------------------------------------------------------- BiCGSTAB kernels (right-preconditioned)
// Shared by solver.cu (raw = 0: the last CTA applies the scalar logic) and dist.cu (raw = 1: the kernel
// only deposits this rank's partial sums in SolveScal::red; after the all-reduce the same logic runs in
// bicg_fin_kernel / inside the peer-memory all-reduce kernel).
__device__ __forceinline__ void bicg_init_fin(SolveScal* s, double tol, double bb) {
    s->rho = bb; s->bb = bb; s->rr = bb;
    s->alpha = 1.0; s->omega = 1.0;
    s->thr = tol < 0 ? -1.0 : tol * tol * bb;
    s->iters = 0;
    s->bad = 0;
    s->breakdown = isfinite(bb) ? 0 : 1;
    s->done = (bb == 0.0 || s->breakdown) ? 1 : 0;
}
// ss = <s,s>; `bad` = the alpha this rank computed was unusable (same on every rank: it derives from all-reduced values)
__device__ __forceinline__ void bicg_s_fin(SolveScal* s, double ss, bool bad) {
    s->rr = ss;   // provisional: ||s||^2
    s->bad = (s->thr >= 0.0 && ss <= s->thr) ? 1 : 0;   // s already below tolerance: take the half step
    if (bad || !isfinite(ss)) { s->breakdown = 1; s->done = 1; }
}
// q0 = <rhat,r>, q1 = <r,r>; alpha, omega and the half-step flag were stored by the x kernel
__device__ __forceinline__ void bicg_x_fin(SolveScal* s, double q0, double q1) {
    const double rho_old = s->rho, alpha = s->alpha, omega = s->omega;
    const bool half = s->half != 0;
    s->rho = q0;
    s->rr  = q1;
    s->iters += 1;
    // beta for the next direction, stored in dots[1] (dots[] are free until the next SpMV)
    const double beta = (q0 / rho_old) * (alpha / omega);
    s->dots[1] = beta;
    if (!isfinite(q0) || !isfinite(q1) || !isfinite(omega)) { s->breakdown = 1; s->done = 1; }
    else if (s->thr >= 0.0 && q1 <= s->thr) s->done = 1;
    else if (half || omega == 0.0 || q0 == 0.0 || !isfinite(beta)) { s->breakdown = 1; s->done = 1; }
}

// r = b ; rhat = b ; p = b ; y = D^-1 p ; x = 0 ; rho = <rhat,r> = bb
static __global__ void __launch_bounds__(EW_THREADS)
bicg_init_kernel(const double* __restrict__ b, const double* __restrict__ dinv, double* __restrict__ x,
                 double* __restrict__ r, double* __restrict__ rhat, double* __restrict__ p, double* __restrict__ y,
                 size_t n, double tol, SolveScal* s, double* partials, uint32_t* ticket, int raw) {
    __shared__ double sh[32];
    double a1 = 0.0;
    for (size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * EW_THREADS) {
        double bi = b[i];
        x[i] = 0.0; r[i] = bi; rhat[i] = bi; p[i] = bi; y[i] = dinv[i] * bi;
        a1 = fma(bi, bi, a1);
    }
    double mine[1];
    mine[0] = abd::block_sum<EW_THREADS>(a1, sh);
    abd::grid_reduce<EW_THREADS, 1>(mine, partials, ticket, sh, [=](const double(&t)[1]) {
        if (raw) { s->red[0] = t[0]; s->red[1] = 0.0; s->done = 0; }
        else bicg_init_fin(s, tol, t[0]);
    });
}
// after v = A y with dots[0] = <v, rhat>: alpha = rho/<rhat,v> ; s = r - alpha v ; z = D^-1 s ; ss = <s,s>
static __global__ void __launch_bounds__(EW_THREADS)
bicg_s_kernel(const double* __restrict__ r, const double* __restrict__ v, const double* __restrict__ dinv,
              double* __restrict__ sv, double* __restrict__ z, size_t n, SolveScal* s, double* partials,
              uint32_t* ticket, int raw) {
    __shared__ double sh[32];
    if (s->done) return;
    const double rv = s->dots[0];
    const bool bad = (rv == 0.0) || !isfinite(rv) || !isfinite(s->rho);
    const double alpha = bad ? 0.0 : s->rho / rv;
    double a0 = 0.0;
    for (size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * EW_THREADS) {
        double si = fma(-alpha, v[i], r[i]);
        sv[i] = si;
        z[i]  = dinv[i] * si;
        a0 = fma(si, si, a0);
    }
    double mine[1];
    mine[0] = abd::block_sum<EW_THREADS>(a0, sh);
    abd::grid_reduce<EW_THREADS, 1>(mine, partials, ticket, sh, [=](const double(&t)[1]) {
        s->alpha = alpha;
        if (raw) { s->red[0] = t[0]; s->red[1] = 0.0; s->half = bad ? 1 : 0; }   // `half` carries `bad` to the fin step
        else bicg_s_fin(s, t[0], bad);
    });
}
// after t = A z with dots = {<t,s>, <t,t>}: omega = ts/tt ; x += alpha y + omega z ; r = s - omega t ;
// rho' = <rhat,r> ; rr = <r,r> ; then beta and the new direction need rho' -> separate kernel.
static __global__ void __launch_bounds__(EW_THREADS)
bicg_x_kernel(const double* __restrict__ y, const double* __restrict__ z, const double* __restrict__ sv,
              const double* __restrict__ t, const double* __restrict__ rhat, double* __restrict__ x,
              double* __restrict__ r, size_t n, SolveScal* s, double* partials, uint32_t* ticket, int raw) {
    __shared__ double sh[32];
    if (s->done) return;
    const double ts = s->dots[0], tt = s->dots[1];
    // ||s|| already tiny (tt == 0): take the half step x += alpha y and stop
    const bool half = (tt == 0.0) || (s->bad != 0);
    const double omega = half ? 0.0 : ts / tt;
    const double alpha = s->alpha;
    double a0 = 0.0, a1 = 0.0;
    for (size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * EW_THREADS) {
        double xi = fma(alpha, y[i], x[i]);
        xi = fma(omega, z[i], xi);
        double ri = fma(-omega, t[i], sv[i]);
        x[i] = xi;
        r[i] = ri;
        a0 = fma(rhat[i], ri, a0);
        a1 = fma(ri, ri, a1);
    }
    double mine[2];
    mine[0] = abd::block_sum<EW_THREADS>(a0, sh);
    mine[1] = abd::block_sum<EW_THREADS>(a1, sh);
    abd::grid_reduce<EW_THREADS, 2>(mine, partials, ticket, sh, [=](const double(&q)[2]) {
        s->omega = omega;
        s->half = half ? 1 : 0;
        if (raw) { s->red[0] = q[0]; s->red[1] = q[1]; }
        else bicg_x_fin(s, q[0], q[1]);
    });
}
// p = r + beta (p - omega v) ; y = D^-1 p
static __global__ void __launch_bounds__(EW_THREADS)
bicg_p_kernel(const double* __restrict__ r, const double* __restrict__ v, const double* __restrict__ dinv,
              double* __restrict__ p, double* __restrict__ y, size_t n, const SolveScal* __restrict__ s) {
    if (s->done) return;
    const double beta = s->dots[1], omega = s->omega;
    for (size_t i = (size_t)blockIdx.x * EW_THREADS + threadIdx.x; i < n; i += (size_t)gridDim.x * EW_THREADS) {
        double pi = fma(beta, fma(-omega, v[i], p[i]), r[i]);
        p[i] = pi;
        y[i] = dinv[i] * pi;
    }
}
// distributed mode: scalar logic after the all-reduce of SolveScal::red.  which: 2 init, 3 after the s kernel, 4 after the x kernel
__device__ __forceinline__ void bicg_fin(SolveScal* s, int which, double tol, double v0, double v1) {
    if (which == 2) bicg_init_fin(s, tol, v0);
    else if (s->done) return;
    else if (which == 3) bicg_s_fin(s, v0, s->half != 0);
    else bicg_x_fin(s, v0, v1);
}

static __global__ void cg_fin_kernel(SolveScal* s, int which, int parity, double tol) {
    if (which == 0) cg_init_fin(s, tol, s->red[0], s->red[1]);
    else if (which >= 2) bicg_fin(s, which, tol, s->red[0], s->red[1]);
    else if (!s->done) cg_update_fin(s, parity, s->red[0], s->red[1], s->bad != 0);
}

// shared host helpers implemented in solver.cu
double* work_vec(Csr* c, int i);
int     read_scal(Csr* c, SolveScal* out);

}  // namespace ab
