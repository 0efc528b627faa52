// stencil.cu — structured-grid matrix generators, built direct-to-CSR on the device, and the
// adjoint of the 2-D variable-coefficient generator.
//
//  * ab_poisson3d_csr: 3-D 7-point Dirichlet Laplacian (diag 6, off -1), rows [row_lo,row_hi) with
//    GLOBAL column ids — the matrix of BASELINE.json configs[1] and configs[2].
//  * ab_poisson3d_kappa_csr: the same pattern with variable coefficients (face weights from a cell field) — the
//    general-CSR leg of the benchmark and of the parity tests (no value dictionary applies).
//  * ab_poisson2d_csr / _update / _grad: the reference's own inverse-problem operator,
//    docs/src/assets/Codes/MPI/ccode/GetPoissonMatrix.h:15-72 (GetPoissonMatrixForward /
//    GetPoissonGrad): off-diagonals kext(nb)+kext(i,j), diagonal -(sum of 4 neighbours + 4 kext(i,j)).
//    Here columns are emitted sorted (CSR), and the scatter-add of GetPoissonGrad is restated as a
//    gather so the result is deterministic.
#include "common.cuh"

namespace ab {

__global__ void p3d_count_kernel(int64_t nx, int64_t ny, int64_t nz, int64_t row_lo, int64_t m, uint32_t* __restrict__ cnt) {
    int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l > m) return;
    if (l == m) { cnt[l] = 0; return; }
    int64_t r = row_lo + l;
    int64_t x = r % nx, y = (r / nx) % ny, z = r / (nx * ny);
    cnt[l] = 1u + (x > 0) + (x < nx - 1) + (y > 0) + (y < ny - 1) + (z > 0) + (z < nz - 1);
}
// kappa == nullptr: constant coefficients (diag 6, off -1).  Otherwise the finite-volume operator of
// -div(kappa grad u) with Dirichlet walls: the face between cells i and j weighs (kappa_i + kappa_j) / 2, a
// wall face weighs kappa_i; off-diagonal = -w, diagonal = sum of the six face weights (symmetric, SPD for
// kappa > 0) — the 3-D analogue of GetPoissonMatrix.h:24-45.  kappa is the GLOBAL cell field [nx*ny*nz].
__global__ void p3d_fill_kernel(int64_t nx, int64_t ny, int64_t nz, int64_t row_lo, int64_t m,
                                const int32_t* __restrict__ rowptr, int32_t* __restrict__ colind,
                                double* __restrict__ values, const double* __restrict__ kappa) {
    int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= m) return;
    int64_t r = row_lo + l;
    int64_t x = r % nx, y = (r / nx) % ny, z = r / (nx * ny);
    int p = rowptr[l];
    const int64_t sxy = nx * ny;
    if (!kappa) {
        if (z > 0)      { colind[p] = (int32_t)(r - sxy); values[p++] = -1.0; }
        if (y > 0)      { colind[p] = (int32_t)(r - nx);  values[p++] = -1.0; }
        if (x > 0)      { colind[p] = (int32_t)(r - 1);   values[p++] = -1.0; }
        colind[p] = (int32_t)r; values[p++] = 6.0;
        if (x < nx - 1) { colind[p] = (int32_t)(r + 1);   values[p++] = -1.0; }
        if (y < ny - 1) { colind[p] = (int32_t)(r + nx);  values[p++] = -1.0; }
        if (z < nz - 1) { colind[p] = (int32_t)(r + sxy); values[p++] = -1.0; }
        return;
    }
    const double kc = kappa[r];
    const double wzm = z > 0      ? 0.5 * (kc + kappa[r - sxy]) : kc;
    const double wym = y > 0      ? 0.5 * (kc + kappa[r - nx])  : kc;
    const double wxm = x > 0      ? 0.5 * (kc + kappa[r - 1])   : kc;
    const double wxp = x < nx - 1 ? 0.5 * (kc + kappa[r + 1])   : kc;
    const double wyp = y < ny - 1 ? 0.5 * (kc + kappa[r + nx])  : kc;
    const double wzp = z < nz - 1 ? 0.5 * (kc + kappa[r + sxy]) : kc;
    if (z > 0)      { colind[p] = (int32_t)(r - sxy); values[p++] = -wzm; }
    if (y > 0)      { colind[p] = (int32_t)(r - nx);  values[p++] = -wym; }
    if (x > 0)      { colind[p] = (int32_t)(r - 1);   values[p++] = -wxm; }
    colind[p] = (int32_t)r; values[p++] = ((wzm + wym) + (wxm + wxp)) + (wyp + wzp);
    if (x < nx - 1) { colind[p] = (int32_t)(r + 1);   values[p++] = -wxp; }
    if (y < ny - 1) { colind[p] = (int32_t)(r + nx);  values[p++] = -wyp; }
    if (z < nz - 1) { colind[p] = (int32_t)(r + sxy); values[p++] = -wzp; }
}

#define KEXT(i, j) kext[(size_t)(i) * (n + 2) + (j)]
__global__ void p2d_count_kernel(int64_t n, uint32_t* __restrict__ cnt) {
    int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l > n * n) return;
    if (l == n * n) { cnt[l] = 0; return; }
    int64_t i = l / n, j = l % n;
    cnt[l] = 1u + (i > 0) + (i < n - 1) + (j > 0) + (j < n - 1);
}
// fill == 1 also writes colind (pattern); values always
__global__ void p2d_fill_kernel(int64_t n, const double* __restrict__ kext, double sign,
                                const int32_t* __restrict__ rowptr, int32_t* __restrict__ colind,
                                double* __restrict__ values, int fill) {
    int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n * n) return;
    const int64_t i0 = l / n, j0 = l % n;   // 0-based grid coordinates
    const int64_t i = i0 + 1, j = j0 + 1;   // position inside the halo-extended array
    const double kc = KEXT(i, j), kn = KEXT(i - 1, j), ks = KEXT(i + 1, j), kw = KEXT(i, j - 1), ke = KEXT(i, j + 1);
    int p = rowptr[l];
    if (i0 > 0)     { if (fill) colind[p] = (int32_t)(l - n); values[p++] = sign * (kn + kc); }
    if (j0 > 0)     { if (fill) colind[p] = (int32_t)(l - 1); values[p++] = sign * (kw + kc); }
    if (fill) colind[p] = (int32_t)l;
    values[p++] = -sign * (ke + kw + ks + kn + 4.0 * kc);
    if (j0 < n - 1) { if (fill) colind[p] = (int32_t)(l + 1); values[p++] = sign * (ke + kc); }
    if (i0 < n - 1) { if (fill) colind[p] = (int32_t)(l + n); values[p++] = sign * (ks + kc); }
}

#define UEXT(i, j) uext[(size_t)(i) * (n + 2) + (j)]
#define XG(i, j) xgrad[(size_t)((i)-1) * n + ((j)-1)]
__global__ void p2d_grad_kernel(int64_t n, const double* __restrict__ xgrad, const double* __restrict__ uext,
                                double sign, double* __restrict__ gk) {
    int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t w = n + 2;
    if (l >= w * w) return;
    const int64_t a = l / w, b = l % w;
    const double ua = UEXT(a, b);
    double g = 0.0;
    // contributions of grid neighbours (i,j) whose stencil reaches (a,b), in the reference's
    // visiting order (row-major over (i,j)): (a-1,b) [its i+1 arm], (a,b-1) [j+1], (a,b) centre,
    // (a,b+1) [j-1], (a+1,b) [i-1]
    const bool row_in = (b >= 1 && b <= n), col_in = (a >= 1 && a <= n);
    if (row_in && a - 1 >= 1 && a - 1 <= n) g += XG(a - 1, b) * (ua - UEXT(a - 1, b));
    if (col_in && b - 1 >= 1 && b - 1 <= n) g += XG(a, b - 1) * (ua - UEXT(a, b - 1));
    if (row_in && col_in)
        g += XG(a, b) * (-4.0 * ua + UEXT(a + 1, b) + UEXT(a - 1, b) + UEXT(a, b + 1) + UEXT(a, b - 1));
    if (col_in && b + 1 >= 1 && b + 1 <= n) g += XG(a, b + 1) * (ua - UEXT(a, b + 1));
    if (row_in && a + 1 >= 1 && a + 1 <= n) g += XG(a + 1, b) * (ua - UEXT(a + 1, b));
    gk[l] = sign * g;
}

static int alloc_csr(Csr* c) {
    AB_TRY(dev_alloc((void**)&c->colind, ((size_t)c->nnz + CSR_PAD) * sizeof(int32_t)));
    AB_TRY(dev_alloc((void**)&c->values, ((size_t)c->nnz + CSR_PAD) * sizeof(double)));
    AB_CUDA(cudaMemsetAsync(c->colind + c->nnz, 0, CSR_PAD * sizeof(int32_t), stream()));
    AB_CUDA(cudaMemsetAsync(c->values + c->nnz, 0, CSR_PAD * sizeof(double), stream()));
    return AB_OK;
}

}  // namespace ab

using namespace ab;

extern "C" {

static int poisson3d_impl(int64_t nx, int64_t ny, int64_t nz, const double* kappa_, int64_t row_lo, int64_t row_hi, int64_t* out_h) {
    AB_TRY(ensure_device());
    if (!out_h || nx < 1 || ny < 1 || nz < 1) return fail(AB_EINVAL, "bad grid");
    const int64_t N = nx * ny * nz;
    if (N >= INT32_MAX) return fail(AB_EUNSUPPORTED, "grid exceeds int32 column ids");
    if (row_lo < 0 || row_hi > N || row_lo > row_hi) return fail(AB_EINVAL, "bad row range");
    std::lock_guard<std::mutex> lk(big_lock());
    const int64_t m = row_hi - row_lo;
    Csr* c = new Csr();
    c->m = m; c->n = N;
    auto body = [&]() -> int {
        In<double> kappa;
        if (kappa_) AB_TRY(kappa.bind(kappa_, (size_t)N));
        DevBuf<uint32_t> cnt, total;
        AB_TRY(cnt.alloc(m + 1)); AB_TRY(total.alloc(1));
        AB_TRY(dev_alloc((void**)&c->rowptr, (size_t)(m + 1) * sizeof(int32_t)));
        unsigned nb = (unsigned)((m + 1 + 255) / 256);
        p3d_count_kernel<<<nb, 256, 0, stream()>>>(nx, ny, nz, row_lo, m, cnt.p);
        AB_LAUNCH_CHECK();
        AB_TRY(scan_exclusive_u32(cnt.p, (uint32_t*)c->rowptr, (size_t)(m + 1), total.p));
        uint32_t nnz = 0;
        AB_CUDA(cudaMemcpyAsync(&nnz, total.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream()));
        AB_CUDA(cudaStreamSynchronize(stream()));
        if (nnz >= (uint32_t)INT32_MAX) return fail(AB_EUNSUPPORTED, "local nnz exceeds int32");
        c->nnz = nnz; c->T = nnz;
        AB_TRY(alloc_csr(c));
        if (m) {
            p3d_fill_kernel<<<(unsigned)((m + 255) / 256), 256, 0, stream()>>>(nx, ny, nz, row_lo, m, c->rowptr, c->colind, c->values,
                                                                             kappa_ ? kappa.d : nullptr);
            AB_LAUNCH_CHECK();
        }
        AB_TRY(csr_finalize(c));
        if (kappa.tmp.p) AB_CUDA(cudaStreamSynchronize(stream()));
        return AB_OK;
    };
    int rc = body();
    if (rc != AB_OK) { delete c; return rc; }
    *out_h = csr_put(c);
    return AB_OK;
}

int ab_poisson3d_csr(int64_t nx, int64_t ny, int64_t nz, int64_t row_lo, int64_t row_hi, int64_t* out_h) {
    return poisson3d_impl(nx, ny, nz, nullptr, row_lo, row_hi, out_h);
}

int ab_poisson3d_kappa_csr(int64_t nx, int64_t ny, int64_t nz, const double* kappa, int64_t row_lo, int64_t row_hi,
                           int64_t* out_h) {
    if (!kappa) return fail(AB_EINVAL, "kappa is null");
    return poisson3d_impl(nx, ny, nz, kappa, row_lo, row_hi, out_h);
}

int ab_poisson2d_csr(int64_t n, const double* kext_, int sign, int64_t* out_h) {
    AB_TRY(ensure_device());
    if (!out_h || !kext_ || n < 1 || (sign != 1 && sign != -1)) return fail(AB_EINVAL, "bad argument");
    if (n * n >= INT32_MAX / 5) return fail(AB_EUNSUPPORTED, "grid too large for int32 nnz");
    std::lock_guard<std::mutex> lk(big_lock());
    const int64_t m = n * n;
    Csr* c = new Csr();
    c->m = m; c->n = m;
    auto body = [&]() -> int {
        In<double> kext;
        AB_TRY(kext.bind(kext_, (size_t)((n + 2) * (n + 2))));
        DevBuf<uint32_t> cnt, total;
        AB_TRY(cnt.alloc(m + 1)); AB_TRY(total.alloc(1));
        AB_TRY(dev_alloc((void**)&c->rowptr, (size_t)(m + 1) * sizeof(int32_t)));
        p2d_count_kernel<<<(unsigned)((m + 1 + 255) / 256), 256, 0, stream()>>>(n, cnt.p);
        AB_LAUNCH_CHECK();
        AB_TRY(scan_exclusive_u32(cnt.p, (uint32_t*)c->rowptr, (size_t)(m + 1), total.p));
        uint32_t nnz = 0;
        AB_CUDA(cudaMemcpyAsync(&nnz, total.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream()));
        AB_CUDA(cudaStreamSynchronize(stream()));
        c->nnz = nnz; c->T = nnz;
        AB_TRY(alloc_csr(c));
        p2d_fill_kernel<<<(unsigned)((m + 255) / 256), 256, 0, stream()>>>(n, kext.d, (double)sign, c->rowptr, c->colind, c->values, 1);
        AB_LAUNCH_CHECK();
        AB_TRY(csr_finalize(c));
        return AB_OK;
    };
    int rc = body();
    if (rc != AB_OK) { delete c; return rc; }
    *out_h = csr_put(c);
    return AB_OK;
}

int ab_poisson2d_update(int64_t h, int64_t n, const double* kext_, int sign) {
    AB_TRY(ensure_device());
    Csr* c = csr_get(h);
    if (!c) return AB_EHANDLE;
    if (!kext_ || c->m != n * n || (sign != 1 && sign != -1)) return fail(AB_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(big_lock());
    In<double> kext;
    AB_TRY(kext.bind(kext_, (size_t)((n + 2) * (n + 2))));
    p2d_fill_kernel<<<(unsigned)((c->m + 255) / 256), 256, 0, stream()>>>(n, kext.d, (double)sign, c->rowptr, c->colind, c->values, 0);
    AB_LAUNCH_CHECK();
    c->dinv_valid = false;
    ++c->values_version;
    if (c->th) { csr_erase(c->th); c->th = 0; dev_free(c->tperm); c->tperm = nullptr; }
    if (kext.tmp.p) AB_CUDA(cudaStreamSynchronize(stream()));
    return AB_OK;
}

int ab_poisson2d_grad(int64_t n, const double* xgrad_, const double* uext_, int sign, double* gk_) {
    AB_TRY(ensure_device());
    if (!xgrad_ || !uext_ || !gk_ || n < 1 || (sign != 1 && sign != -1)) return fail(AB_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(big_lock());
    const size_t w2 = (size_t)((n + 2) * (n + 2));
    In<double> xg, ue; Out<double> gk;
    AB_TRY(xg.bind(xgrad_, (size_t)(n * n))); AB_TRY(ue.bind(uext_, w2)); AB_TRY(gk.bind(gk_, w2));
    p2d_grad_kernel<<<(unsigned)((w2 + 255) / 256), 256, 0, stream()>>>(n, xg.d, ue.d, (double)sign, gk.d);
    AB_LAUNCH_CHECK();
    AB_TRY(gk.commit());
    if (xg.tmp.p || ue.tmp.p) AB_CUDA(cudaStreamSynchronize(stream()));
    return AB_OK;
}

}  // extern "C"
