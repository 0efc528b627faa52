// spmm.cu — SpMM Y = A X (k right-hand sides, row-major like TensorFlow), its two gradients,
// and the public ab_spmm* entry points (k == 1 routes to the SpMV kernels in spmv.cu).
//
// Reference: src/sparse.jl:227-264 -> tf.sparse.sparse_dense_matmul and its registered gradient
// (TF 1.15 _SparseTensorDenseMatMulGrad; restated in oracle/adcme_oracle.c):
//     Y[i,:]  = sum_k v_k X[j_k,:]
//     dX      = A' dY
//     dv[k]   = <dY[i_k,:], X[j_k,:]>      one value per INPUT triplet
#include "common.cuh"

namespace ab {

// One group of CL lanes per row; lane c of the group owns columns c, c+CL, ...  (NC of them).
template <int CL, int NC>
__global__ void __launch_bounds__(256)
spmm_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
            const double* __restrict__ values, const double* __restrict__ X, double* __restrict__ Y,
            int64_t m, int k, int c0) {
    constexpr int RPB = 256 / CL;
    const int sub = threadIdx.x % CL;
    for (int64_t rb = (int64_t)blockIdx.x * RPB; rb < m; rb += (int64_t)gridDim.x * RPB) {
        const int64_t row = rb + threadIdx.x / CL;
        if (row >= m) continue;
        double acc[NC];
#pragma unroll
        for (int q = 0; q < NC; ++q) acc[q] = 0.0;
        const int a = rowptr[row], b = rowptr[row + 1];
        for (int j = a; j < b; ++j) {
            const double v = values[j];
            const double* xr = X + (size_t)colind[j] * k + c0;
#pragma unroll
            for (int q = 0; q < NC; ++q) {
                const int cc = sub + q * CL;
                if (c0 + cc < k) acc[q] = fma(v, __ldg(xr + cc), acc[q]);
            }
        }
        double* yr = Y + (size_t)row * k + c0;
#pragma unroll
        for (int q = 0; q < NC; ++q) {
            const int cc = sub + q * CL;
            if (c0 + cc < k) yr[cc] = acc[q];
        }
    }
}

// k >= 32: one warp per row, entries fetched 32 at a time (coalesced) and broadcast by shuffle.
template <int NC>
__global__ void __launch_bounds__(256)
spmm_warp_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                 const double* __restrict__ values, const double* __restrict__ X, double* __restrict__ Y,
                 int64_t m, int k, int c0) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * 256) >> 5;
    for (int64_t row = warp0; row < m; row += nwarp) {
        double acc[NC];
#pragma unroll
        for (int q = 0; q < NC; ++q) acc[q] = 0.0;
        const int a = rowptr[row], b = rowptr[row + 1];
        for (int base = a; base < b; base += 32) {
            const int cnt = min(32, b - base);
            int cj = 0;
            double vj = 0.0;
            if (lane < cnt) { cj = colind[base + lane]; vj = values[base + lane]; }
            for (int q = 0; q < cnt; ++q) {
                const int cq = __shfl_sync(0xffffffffu, cj, q);
                const double vq = __shfl_sync(0xffffffffu, vj, q);
                const double* xr = X + (size_t)cq * k + c0;
#pragma unroll
                for (int u = 0; u < NC; ++u) {
                    const int cc = lane + 32 * u;
                    if (c0 + cc < k) acc[u] = fma(vq, __ldg(xr + cc), acc[u]);
                }
            }
        }
        double* yr = Y + (size_t)row * k + c0;
#pragma unroll
        for (int u = 0; u < NC; ++u) {
            const int cc = lane + 32 * u;
            if (c0 + cc < k) yr[cc] = acc[u];
        }
    }
}

int launch_spmm(const Csr* c, const double* X, int64_t k, double* Y) {
    if (c->m == 0 || k == 0) return AB_OK;
    if (k == 1) return launch_spmv(c, X, Y, nullptr, nullptr, nullptr, nullptr, nullptr);
    if (k > INT32_MAX / 2) return fail(AB_EUNSUPPORTED, "k too large");
    const int kk = (int)k;
    if (kk >= 32) {
        int64_t nblk = (c->m + 7) / 8;
        int64_t cap  = (int64_t)sm_count() * 8 * 8;
        unsigned grid = (unsigned)(nblk < cap ? nblk : cap);
        for (int c0 = 0; c0 < kk; c0 += 128) {
            int rem = kk - c0;
            if (rem > 96) spmm_warp_kernel<4><<<grid, 256, 0, stream()>>>(c->rowptr, c->colind, c->values, X, Y, c->m, kk, c0);
            else if (rem > 64) spmm_warp_kernel<3><<<grid, 256, 0, stream()>>>(c->rowptr, c->colind, c->values, X, Y, c->m, kk, c0);
            else if (rem > 32) spmm_warp_kernel<2><<<grid, 256, 0, stream()>>>(c->rowptr, c->colind, c->values, X, Y, c->m, kk, c0);
            else spmm_warp_kernel<1><<<grid, 256, 0, stream()>>>(c->rowptr, c->colind, c->values, X, Y, c->m, kk, c0);
            AB_LAUNCH_CHECK();
        }
        return AB_OK;
    }
    int CL = 2;
    while (CL < kk) CL <<= 1;   // 2,4,8,16,32
    const int64_t rpb = 256 / CL;
    int64_t nblk = (c->m + rpb - 1) / rpb;
    int64_t cap  = (int64_t)sm_count() * 8 * 8;
    unsigned grid = (unsigned)(nblk < cap ? nblk : cap);
#define AB_SPMM(CLL) spmm_kernel<CLL, 1><<<grid, 256, 0, stream()>>>(c->rowptr, c->colind, c->values, X, Y, c->m, kk, 0)
    switch (CL) {
        case 2: AB_SPMM(2); break;
        case 4: AB_SPMM(4); break;
        case 8: AB_SPMM(8); break;
        case 16: AB_SPMM(16); break;
        default: AB_SPMM(32); break;
    }
#undef AB_SPMM
    AB_LAUNCH_CHECK();
    return AB_OK;
}

// SDDMM sampled at the stored pattern: dv[e] = <dY[row_e,:], X[col_e,:]>, CL lanes per entry
template <int CL>
__global__ void __launch_bounds__(256)
sddmm_kernel(const int32_t* __restrict__ entry_row, const int32_t* __restrict__ colind,
             const double* __restrict__ dY, const double* __restrict__ X, size_t nnz, int k,
             double* __restrict__ dv) {
    constexpr int EPB = 256 / CL;
    const int sub = threadIdx.x % CL;
    for (size_t eb = (size_t)blockIdx.x * EPB; eb < nnz; eb += (size_t)gridDim.x * EPB) {
        const size_t e = eb + threadIdx.x / CL;
        double acc = 0.0;
        if (e < nnz) {
            const double* yr = dY + (size_t)entry_row[e] * k;
            const double* xr = X + (size_t)colind[e] * k;
            for (int cc = sub; cc < k; cc += CL) acc = fma(__ldg(yr + cc), __ldg(xr + cc), acc);
        }
#pragma unroll
        for (int o = CL / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (e < nnz && sub == 0) dv[e] = acc;
    }
}

__global__ void gather_slot_kernel(const int32_t* __restrict__ slot, const double* __restrict__ dv, size_t T,
                                   double* __restrict__ out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T) out[t] = dv[slot[t]];
}

static int launch_sddmm(Csr* c, const double* dY, const double* X, int k, double* dv) {
    if (c->nnz == 0) return AB_OK;
    AB_TRY(csr_ensure_entry_row(c));
    int CL = 1;
    while (CL < k && CL < 32) CL <<= 1;
    const size_t epb = 256 / CL;
    size_t nblk = ((size_t)c->nnz + epb - 1) / epb;
    size_t cap  = (size_t)sm_count() * 8 * 16;
    unsigned grid = (unsigned)(nblk < cap ? nblk : cap);
#define AB_SD(CLL) sddmm_kernel<CLL><<<grid, 256, 0, stream()>>>(c->entry_row, c->colind, dY, X, (size_t)c->nnz, k, dv)
    switch (CL) {
        case 1: AB_SD(1); break;
        case 2: AB_SD(2); break;
        case 4: AB_SD(4); break;
        case 8: AB_SD(8); break;
        case 16: AB_SD(16); break;
        default: AB_SD(32); break;
    }
#undef AB_SD
    AB_LAUNCH_CHECK();
    return AB_OK;
}

}  // namespace ab

using namespace ab;

extern "C" {

int ab_spmm(int64_t h, int transA, const double* X_, int64_t k, double* Y_) {
    AB_TRY(ensure_device());
    Csr* c = csr_get(h);
    if (!c) return AB_EHANDLE;
    if (k < 0) return fail(AB_EINVAL, "k < 0");
    std::lock_guard<std::mutex> lk(big_lock());
    Csr* a = c;
    if (transA) AB_TRY(csr_ensure_transpose(c, &a));
    if ((a->n * k > 0 && !X_) || (a->m * k > 0 && !Y_)) return fail(AB_EINVAL, "null operand");
    In<double> X; Out<double> Y;
    AB_TRY(X.bind(X_, (size_t)(a->n * k))); AB_TRY(Y.bind(Y_, (size_t)(a->m * k)));
    AB_TRY(launch_spmm(a, X.d, k, Y.d));
    AB_TRY(Y.commit());
    if (X.tmp.p) AB_CUDA(cudaStreamSynchronize(stream()));
    return AB_OK;
}

int ab_spmm_grad_dense(int64_t h, const double* dY, int64_t k, double* dX) {
    return ab_spmm(h, 1, dY, k, dX);
}

int ab_spmm_grad_values(int64_t h, const double* dY_, const double* X_, int64_t k, double* dvv_) {
    AB_TRY(ensure_device());
    Csr* c = csr_get(h);
    if (!c) return AB_EHANDLE;
    if (k < 0 || k > INT32_MAX / 2) return fail(AB_EINVAL, "bad k");
    if (c->T == 0) return AB_OK;
    if (!dY_ || !X_ || !dvv_) return fail(AB_EINVAL, "null operand");
    std::lock_guard<std::mutex> lk(big_lock());
    In<double> dY, X; Out<double> dvv;
    AB_TRY(dY.bind(dY_, (size_t)(c->m * k))); AB_TRY(X.bind(X_, (size_t)(c->n * k)));
    AB_TRY(dvv.bind(dvv_, (size_t)c->T));
    if (c->slot) {
        DevBuf<double> dv;
        AB_TRY(dv.alloc((size_t)c->nnz));
        AB_TRY(launch_sddmm(c, dY.d, X.d, (int)k, dv.p));
        gather_slot_kernel<<<(unsigned)((c->T + 255) / 256), 256, 0, stream()>>>(c->slot, dv.p, (size_t)c->T, dvv.d);
        AB_LAUNCH_CHECK();
        AB_TRY(dvv.commit());
        AB_CUDA(cudaStreamSynchronize(stream()));
    } else {
        AB_TRY(launch_sddmm(c, dY.d, X.d, (int)k, dvv.d));
        AB_TRY(dvv.commit());
        if (dY.tmp.p || X.tmp.p) AB_CUDA(cudaStreamSynchronize(stream()));
    }
    return AB_OK;
}

// ---- measurement hooks ---------------------------------------------------------------
int ab_bench_spmv(int64_t h, const double* x, double* y, int reps, int with_dot, double* ms) {
    AB_TRY(ensure_device());
    Csr* c = csr_get(h);
    if (!c) return AB_EHANDLE;
    if (!x || !y || !ms || reps < 1) return fail(AB_EINVAL, "bad argument");
    if (!is_device_ptr(x) || !is_device_ptr(y)) return fail(AB_EINVAL, "ab_bench_spmv needs device vectors");
    std::lock_guard<std::mutex> lk(big_lock());
    AB_TRY(csr_ensure_work(c, 8));
    double* dots = (double*)c->scal + 2;   // SolveScal::dots
    cudaEvent_t e0, e1;
    AB_CUDA(cudaEventCreate(&e0)); AB_CUDA(cudaEventCreate(&e1));
    AB_CUDA(cudaEventRecord(e0, stream()));
    for (int i = 0; i < reps; ++i)
        AB_TRY(launch_spmv(c, x, y, with_dot ? x : nullptr, dots, c->partials, c->ticket, nullptr));
    AB_CUDA(cudaEventRecord(e1, stream()));
    AB_CUDA(cudaEventSynchronize(e1));
    float t = 0;
    AB_CUDA(cudaEventElapsedTime(&t, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *ms = (double)t / reps;
    return AB_OK;
}

}  // extern "C"
