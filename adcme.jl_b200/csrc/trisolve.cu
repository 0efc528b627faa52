// trisolve.cu — tridiagonal solve `trisolve(a, b, c, d)` and its gradient.
//
// Reference: src/sparse.jl:731-739 -> TriSolve op, deps/CustomOps/TriSolve/TriSolve.h:1-37 (Thomas
// algorithm on one CPU thread; backward = the transposed system + three element-wise products).
// a[n-1] sub-diagonal (a[i-1] multiplies x[i-1] in row i), b[n] diagonal, c[n-1] super-diagonal, d[n] rhs.
//
// The Thomas recurrence is a length-n dependency chain; the device form is parallel cyclic reduction:
// ceil(log2 n) sweeps, each eliminating the couplings at distance s and doubling s, every row handled by
// one thread with coalesced reads at i-s, i, i+s — HBM-bound streaming work (64 B per row per sweep).
// Stable for the diagonally dominant systems implicit 1-D schemes produce; results agree with Thomas to
// rounding (parity bar in the tests: 1e-10 relative; the reference's own test asserts 1e-3).
#include "common.cuh"

namespace ab {

// one PCR sweep at distance s; neighbours outside [0,n) act as identity rows (a = c = d = 0, b = 1)
__global__ void __launch_bounds__(256)
pcr_sweep_kernel(const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ c,
                 const double* __restrict__ d, double* __restrict__ na, double* __restrict__ nb, double* __restrict__ nc,
                 double* __restrict__ nd, int64_t n, int64_t s) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double ai = a[i], bi = b[i], ci = c[i], di = d[i];
    double am = 0.0, bm = 1.0, cm = 0.0, dm = 0.0, ap = 0.0, bp = 1.0, cp = 0.0, dp = 0.0;
    const int64_t im = i - s, ip = i + s;
    if (im >= 0) { am = a[im]; bm = b[im]; cm = c[im]; dm = d[im]; }
    if (ip < n)  { ap = a[ip]; bp = b[ip]; cp = c[ip]; dp = d[ip]; }
    const double k1 = ai / bm, k2 = ci / bp;
    na[i] = -am * k1;
    nb[i] = bi - cm * k1 - ap * k2;
    nc[i] = -cp * k2;
    nd[i] = di - dm * k1 - dp * k2;
}
// row-indexed copies: la[i] = sub-diagonal entry of row i (0 for row 0), lc[i] = super-diagonal (0 for the last row)
__global__ void tri_pack_kernel(const double* __restrict__ a, const double* __restrict__ b, const double* __restrict__ c,
                                const double* __restrict__ d, int64_t n, double* __restrict__ la, double* __restrict__ lb,
                                double* __restrict__ lc, double* __restrict__ ld) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    la[i] = i > 0 ? a[i - 1] : 0.0;
    lb[i] = b[i];
    lc[i] = i < n - 1 ? c[i] : 0.0;
    ld[i] = d[i];
}
__global__ void tri_finish_kernel(const double* __restrict__ b, const double* __restrict__ d, int64_t n, double* __restrict__ x,
                                  int* __restrict__ bad) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = d[i] / b[i];
    if (!isfinite(v)) atomicExch(bad, 1);
    x[i] = v;
}
// TriSolve_backward, TriSolve.h:31-36
__global__ void tri_grad_kernel(const double* __restrict__ gd, const double* __restrict__ x, int64_t n,
                                double* __restrict__ ga, double* __restrict__ gb, double* __restrict__ gc) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double g = gd[i];
    if (ga && i > 0) ga[i - 1] = -g * x[i - 1];
    if (gb) gb[i] = -g * x[i];
    if (gc && i < n - 1) gc[i] = -g * x[i + 1];
}

// device pointers in, device x out
static int tri_solve_dev(int64_t n, const double* a, const double* b, const double* c, const double* d, double* x) {
    DevBuf<double> w[8];
    for (auto& v : w) AB_TRY(v.alloc((size_t)n));
    DevBuf<int> bad;
    AB_TRY(bad.alloc(1));
    AB_CUDA(cudaMemsetAsync(bad.p, 0, sizeof(int), stream()));
    const unsigned nb = (unsigned)((n + 255) / 256);
    tri_pack_kernel<<<nb, 256, 0, stream()>>>(a, b, c, d, n, w[0].p, w[1].p, w[2].p, w[3].p);
    AB_LAUNCH_CHECK();
    int cur = 0;
    for (int64_t s = 1; s < n; s <<= 1) {
        const int o = cur ? 0 : 4, q = cur ? 4 : 0;
        pcr_sweep_kernel<<<nb, 256, 0, stream()>>>(w[q].p, w[q + 1].p, w[q + 2].p, w[q + 3].p, w[o].p, w[o + 1].p, w[o + 2].p,
                                                  w[o + 3].p, n, s);
        AB_LAUNCH_CHECK();
        cur ^= 1;
    }
    const int q = cur ? 4 : 0;
    tri_finish_kernel<<<nb, 256, 0, stream()>>>(w[q + 1].p, w[q + 3].p, n, x, bad.p);
    AB_LAUNCH_CHECK();
    int hb = 0;
    AB_CUDA(cudaMemcpyAsync(&hb, bad.p, sizeof(int), cudaMemcpyDeviceToHost, stream()));
    AB_CUDA(cudaStreamSynchronize(stream()));
    if (hb) return fail(AB_EBREAKDOWN, "trisolve: zero pivot or non-finite value (the reduction needs a regular, diagonally dominant system)");
    return AB_OK;
}

}  // namespace ab

using namespace ab;

extern "C" {

int ab_tri_solve(int64_t n, const double* a_, const double* b_, const double* c_, const double* d_, double* x_) {
    AB_TRY(ensure_device());
    if (n < 0) return fail(AB_EINVAL, "negative size");
    if (n == 0) return AB_OK;
    if (!b_ || !d_ || !x_ || (n > 1 && (!a_ || !c_))) return fail(AB_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(big_lock());
    In<double> a, b, c, d; Out<double> x;
    AB_TRY(a.bind(a_, (size_t)(n - 1))); AB_TRY(b.bind(b_, (size_t)n)); AB_TRY(c.bind(c_, (size_t)(n - 1)));
    AB_TRY(d.bind(d_, (size_t)n)); AB_TRY(x.bind(x_, (size_t)n));
    AB_TRY(tri_solve_dev(n, a.d, b.d, c.d, d.d, x.d));
    AB_TRY(x.commit());
    AB_CUDA(cudaStreamSynchronize(stream()));
    return AB_OK;
}

// TriSolveGrad(grad_x, x, a, b, c, d) -> (grad_a, grad_b, grad_c, grad_d)   (TriSolve.cpp:35-45)
int ab_tri_solve_grad(int64_t n, const double* grad_x_, const double* x_, const double* a_, const double* b_, const double* c_,
                      double* grad_a_, double* grad_b_, double* grad_c_, double* grad_d_) {
    AB_TRY(ensure_device());
    if (n < 0) return fail(AB_EINVAL, "negative size");
    if (n == 0) return AB_OK;
    if (!grad_x_ || !x_ || !b_ || (n > 1 && (!a_ || !c_))) return fail(AB_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(big_lock());
    In<double> gx, x, a, b, c; Out<double> ga, gb, gc, gd;
    AB_TRY(gx.bind(grad_x_, (size_t)n)); AB_TRY(x.bind(x_, (size_t)n));
    AB_TRY(a.bind(a_, (size_t)(n - 1))); AB_TRY(b.bind(b_, (size_t)n)); AB_TRY(c.bind(c_, (size_t)(n - 1)));
    AB_TRY(ga.bind(grad_a_, (size_t)(n - 1))); AB_TRY(gb.bind(grad_b_, (size_t)n)); AB_TRY(gc.bind(grad_c_, (size_t)(n - 1)));
    DevBuf<double> gdtmp;
    double* gdp = nullptr;
    if (grad_d_) { AB_TRY(gd.bind(grad_d_, (size_t)n)); gdp = gd.d; } else { AB_TRY(gdtmp.alloc((size_t)n)); gdp = gdtmp.p; }
    AB_TRY(tri_solve_dev(n, c.d, b.d, a.d, gx.d, gdp));          // the transposed system: sub <-> super (TriSolve.h:31)
    tri_grad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream()>>>(gdp, x.d, n, ga.d, gb.d, gc.d);
    AB_LAUNCH_CHECK();
    AB_TRY(ga.commit()); AB_TRY(gb.commit()); AB_TRY(gc.commit()); AB_TRY(gd.commit());
    AB_CUDA(cudaStreamSynchronize(stream()));
    return AB_OK;
}

}  // extern "C"
