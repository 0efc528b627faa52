// spmm.cu — SpMM Y = A X (k right-hand sides, row-major like TensorFlow), its two gradients,
// and the public ab_spmm* entry points (k == 1 routes to the SpMV kernels in spmv.cu).
//
// Reference: src/sparse.jl:227-264 -> tf.sparse.sparse_dense_matmul and its registered gradient
// (TF 1.15 _SparseTensorDenseMatMulGrad; restated in oracle/adcme_oracle.c):
//     Y[i,:]  = sum_k v_k X[j_k,:]
//     dX      = A' dY
//     dv[k]   = <dY[i_k,:], X[j_k,:]>      one value per INPUT triplet
//
// k >= 32: one warp per row; the row's entries are fetched 32 at a time (coalesced) and broadcast
// by shuffle; lane c owns columns c, c+32, … so every X-row gather is one or more full 256-byte
// requests, UB of them in flight per warp.  Rows longer than LONG_ROW (power-law tails) are cut into
// CHUNK-entry pieces handled by separate warps, with a fixed-order combine — no atomics, so the
// result is deterministic.  Algorithmic bytes: 12*nnz + 4*(m+1) + 8*k*(n+m) compulsory;
// 12*nnz + 8*k*(nnz+m) under the no-reuse gather model (DESIGN.md reports both).
#include "common.cuh"

namespace ab {

constexpr int LONG_ROW = 4096;
constexpr int CHUNK    = 1024;

// ---------------------------------------------------------------- long-row work list
__global__ void long_flag_kernel(const int32_t* __restrict__ rowptr, int64_t m, uint32_t* __restrict__ flags) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < m) flags[r] = (rowptr[r + 1] - rowptr[r] > LONG_ROW) ? 1u : 0u;
}
__global__ void long_compact_kernel(const int32_t* __restrict__ rowptr, int64_t m, const uint32_t* __restrict__ excl,
                                    int32_t* __restrict__ rows, uint32_t* __restrict__ nchunk) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    int len = rowptr[r + 1] - rowptr[r];
    if (len > LONG_ROW) { rows[excl[r]] = (int32_t)r; nchunk[excl[r]] = (uint32_t)((len + CHUNK - 1) / CHUNK); }
}
__global__ void long_fill_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ rows,
                                 const int32_t* __restrict__ chunk0, int64_t nlong, int32_t* __restrict__ crow,
                                 int32_t* __restrict__ cstart) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlong) return;
    const int r = rows[i], a = rowptr[r], b = rowptr[r + 1];
    int cidx = chunk0[i];
    for (int s = a; s < b; s += CHUNK, ++cidx) { crow[cidx] = r; cstart[cidx] = s; }
}

static int csr_ensure_long(Csr* c) {
    if (c->long_nrows >= 0) return AB_OK;
    if (c->max_row_nnz <= LONG_ROW || c->m == 0) { c->long_nrows = 0; c->long_nchunks = 0; return AB_OK; }
    DevBuf<uint32_t> flags, total, nchunk;
    AB_TRY(flags.alloc(c->m)); AB_TRY(total.alloc(1));
    unsigned nb = (unsigned)((c->m + 255) / 256);
    long_flag_kernel<<<nb, 256, 0, stream()>>>(c->rowptr, c->m, flags.p);
    AB_LAUNCH_CHECK();
    AB_TRY(scan_exclusive_u32(flags.p, flags.p, (size_t)c->m, total.p));
    uint32_t nlong = 0;
    AB_CUDA(cudaMemcpyAsync(&nlong, total.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream()));
    AB_CUDA(cudaStreamSynchronize(stream()));
    AB_TRY(dev_alloc((void**)&c->long_rows, (size_t)(nlong + 1) * sizeof(int32_t)));
    AB_TRY(dev_alloc((void**)&c->long_row_chunk0, (size_t)(nlong + 1) * sizeof(int32_t)));
    AB_TRY(nchunk.alloc(nlong + 1));
    AB_CUDA(cudaMemsetAsync(nchunk.p, 0, (size_t)(nlong + 1) * sizeof(uint32_t), stream()));
    long_compact_kernel<<<nb, 256, 0, stream()>>>(c->rowptr, c->m, flags.p, c->long_rows, nchunk.p);
    AB_LAUNCH_CHECK();
    AB_TRY(scan_exclusive_u32(nchunk.p, (uint32_t*)c->long_row_chunk0, (size_t)nlong + 1, total.p));
    uint32_t nch = 0;
    AB_CUDA(cudaMemcpyAsync(&nch, total.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream()));
    AB_CUDA(cudaStreamSynchronize(stream()));
    AB_TRY(dev_alloc((void**)&c->long_chunk_row, (size_t)(nch + 1) * sizeof(int32_t)));
    AB_TRY(dev_alloc((void**)&c->long_chunk_start, (size_t)(nch + 1) * sizeof(int32_t)));
    long_fill_kernel<<<(nlong + 255) / 256, 256, 0, stream()>>>(c->rowptr, c->long_rows, c->long_row_chunk0, nlong,
                                                               c->long_chunk_row, c->long_chunk_start);
    AB_LAUNCH_CHECK();
    c->long_nrows = nlong; c->long_nchunks = nch;
    return AB_OK;
}

// ---------------------------------------------------------------- SpMM
// One group of CL lanes per row; lane c of the group owns column c (k <= 32 here).
template <int CL>
__global__ void __launch_bounds__(256)
spmm_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
            const double* __restrict__ values, const double* __restrict__ X, double* __restrict__ Y,
            int64_t m, int k) {
    constexpr int RPB = 256 / CL;
    const int sub = threadIdx.x % CL;
    for (int64_t rb = (int64_t)blockIdx.x * RPB; rb < m; rb += (int64_t)gridDim.x * RPB) {
        const int64_t row = rb + threadIdx.x / CL;
        if (row >= m || sub >= k) continue;
        double acc = 0.0;
        const int a = rowptr[row], b = rowptr[row + 1];
        int j = a;
        for (; j + 4 <= b; j += 4) {
            double v[4], xv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { v[u] = values[j + u]; xv[u] = __ldg(X + (size_t)colind[j + u] * k + sub); }
#pragma unroll
            for (int u = 0; u < 4; ++u) acc = fma(v[u], xv[u], acc);
        }
        for (; j < b; ++j) acc = fma(values[j], __ldg(X + (size_t)colind[j] * k + sub), acc);
        Y[(size_t)row * k + sub] = acc;
    }
}

// accumulates entries [a,b) of one row into acc[] (lane owns columns c0+lane+32u)
template <int NC>
__device__ __forceinline__ void spmm_range(const int32_t* __restrict__ colind, const double* __restrict__ values,
                                           const double* __restrict__ X, int k, int c0, int a, int b, int lane,
                                           double (&acc)[NC]) {
    constexpr int UB = NC == 1 ? 8 : 4;
    for (int base = a; base < b; base += 32) {
        const int cnt = min(32, b - base);
        int cj = 0;
        double vj = 0.0;
        if (lane < cnt) { cj = colind[base + lane]; vj = values[base + lane]; }
        // UB entries at a time: all their X-row gathers are issued before the first FMA needs one
        for (int q0 = 0; q0 < cnt; q0 += UB) {
            double xv[UB][NC], vq[UB];
#pragma unroll
            for (int b8 = 0; b8 < UB; ++b8) {
                const int q = q0 + b8;
                const int cq = __shfl_sync(0xffffffffu, cj, q & 31);
                vq[b8] = __shfl_sync(0xffffffffu, vj, q & 31);
                const double* xr = X + (size_t)cq * k + c0;
#pragma unroll
                for (int u = 0; u < NC; ++u) {
                    const int cc = lane + 32 * u;
                    xv[b8][u] = (q < cnt && c0 + cc < k) ? __ldg(xr + cc) : 0.0;
                }
            }
#pragma unroll
            for (int b8 = 0; b8 < UB; ++b8)
#pragma unroll
                for (int u = 0; u < NC; ++u) acc[u] = fma(vq[b8], xv[b8][u], acc[u]);
        }
    }
}

// k >= 32: one warp per row (rows longer than long_thresh are left to the chunk kernels)
template <int NC>
__global__ void __launch_bounds__(256)
spmm_warp_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                 const double* __restrict__ values, const double* __restrict__ X, double* __restrict__ Y,
                 int64_t m, int k, int c0, int long_thresh) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * 256) >> 5;
    for (int64_t row = warp0; row < m; row += nwarp) {
        const int a = rowptr[row], b = rowptr[row + 1];
        if (b - a > long_thresh) continue;
        double acc[NC];
#pragma unroll
        for (int q = 0; q < NC; ++q) acc[q] = 0.0;
        spmm_range<NC>(colind, values, X, k, c0, a, b, lane, acc);
        double* yr = Y + (size_t)row * k + c0;
#pragma unroll
        for (int u = 0; u < NC; ++u) {
            const int cc = lane + 32 * u;
            if (c0 + cc < k) yr[cc] = acc[u];
        }
    }
}
// k >= 32, short rows: one warp per GROUP of 8 consecutive rows.  A power-law matrix is mostly rows of 1-3
// entries; with a warp per row such a row keeps one 256-byte X gather in flight and the kernel is latency-bound
// (C5: 0.51 of the gather-model roofline).  Here the group's entries (<= 32 in total) are fetched by one
// coalesced load, their X-row gathers are issued eight at a time ACROSS row boundaries, and the running sum is
// flushed to Y whenever the row changes (warp-uniform branch).  Measured alternatives at C5 x 0.4 (25.7 ms with this
// kernel): 16 gathers in flight per warp (80 registers or spills) 30.6-39.6 ms; groups of 24 rows 25.1 ms; entry-balanced
// tasks (the rows starting in one 64-entry block of the CSR arrays) 26.3 ms — the row structure is not what limits it.  Per-row summation order is unchanged (stored
// order), so Y is bit-identical to the warp-per-row kernel.  A group with more than 32 entries is walked in
// 32-entry windows by the same code; only a group that contains a long row (chunk kernels) goes row by row.
constexpr int SPMM_GROUP = 8;
template <int NC>
__global__ void __launch_bounds__(256, 4)
spmm_group_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                  const double* __restrict__ values, const double* __restrict__ X, double* __restrict__ Y,
                  int64_t m, int k, int c0, int long_thresh) {
    constexpr int UB = NC == 1 ? 8 : 4;     // measured: 16 in flight (80 registers, spills) 39.6 ms vs 25.7 ms at C5 x 0.4
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * 256) >> 5;
    const int64_t ngroups = (m + SPMM_GROUP - 1) / SPMM_GROUP;
    for (int64_t grp = warp0; grp < ngroups; grp += nwarp) {
        const int64_t row0 = grp * SPMM_GROUP;
        const int nr = (int)min((int64_t)SPMM_GROUP, m - row0);
        const int rp = lane <= nr ? rowptr[row0 + lane] : 0;
        const int a = __shfl_sync(0xffffffffu, rp, 0), b = __shfl_sync(0xffffffffu, rp, nr);
        const int cnt = b - a;
        const int rp_next = __shfl_down_sync(0xffffffffu, rp, 1);
        const int mylen = lane < nr ? rp_next - rp : 0;
        const bool has_long = __any_sync(0xffffffffu, lane < nr && mylen > long_thresh);
        if (has_long) {
            // a long row sits in this group: row at a time (the long row itself is left to the chunk kernels)
            for (int rr = 0; rr < nr; ++rr) {
                const int ra = __shfl_sync(0xffffffffu, rp, rr), rb = __shfl_sync(0xffffffffu, rp, rr + 1);
                if (rb - ra > long_thresh) continue;
                double acc[NC];
#pragma unroll
                for (int q = 0; q < NC; ++q) acc[q] = 0.0;
                spmm_range<NC>(colind, values, X, k, c0, ra, rb, lane, acc);
                double* yr = Y + (size_t)(row0 + rr) * k + c0;
#pragma unroll
                for (int u = 0; u < NC; ++u) {
                    const int cc = lane + 32 * u;
                    if (c0 + cc < k) yr[cc] = acc[u];
                }
            }
            continue;
        }
        double acc[NC];
#pragma unroll
        for (int q = 0; q < NC; ++q) acc[q] = 0.0;
        int cur = 0;
        auto flush_to = [&](int upto) {
            while (cur < upto) {
                double* yr = Y + (size_t)(row0 + cur) * k + c0;
#pragma unroll
                for (int u = 0; u < NC; ++u) {
                    const int cc = lane + 32 * u;
                    if (c0 + cc < k) yr[cc] = acc[u];
                    acc[u] = 0.0;
                }
                ++cur;
            }
        };
        // the group's entries in windows of 32 (one coalesced index/value load each); the X-row gathers run UB at a time
        // straight across row boundaries AND window boundaries of the group
        for (int w0 = 0; w0 < cnt; w0 += 32) {
            const int wn = min(32, cnt - w0);
            int cj = 0, rid = 0;
            double vj = 0.0;
            if (lane < wn) { cj = colind[a + w0 + lane]; vj = values[a + w0 + lane]; }
            for (int rr = 1; rr <= nr; ++rr) rid += (__shfl_sync(0xffffffffu, rp, rr) <= a + w0 + lane) ? 1 : 0;   // row (within the group) of entry `lane`
            for (int q0 = 0; q0 < wn; q0 += UB) {
                double xv[UB][NC], vq[UB];
                int rq[UB];
#pragma unroll
                for (int b8 = 0; b8 < UB; ++b8) {
                    const int q = q0 + b8;
                    const int cq = __shfl_sync(0xffffffffu, cj, q & 31);
                    vq[b8] = __shfl_sync(0xffffffffu, vj, q & 31);
                    rq[b8] = __shfl_sync(0xffffffffu, rid, q & 31);
                    const double* xr = X + (size_t)cq * k + c0;
#pragma unroll
                    for (int u = 0; u < NC; ++u) {
                        const int cc = lane + 32 * u;
                        xv[b8][u] = (q < wn && c0 + cc < k) ? __ldg(xr + cc) : 0.0;
                    }
                }
#pragma unroll
                for (int b8 = 0; b8 < UB; ++b8) {
                    if (q0 + b8 < wn) {
                        flush_to(rq[b8]);
#pragma unroll
                        for (int u = 0; u < NC; ++u) acc[u] = fma(vq[b8], xv[b8][u], acc[u]);
                    }
                }
            }
        }
        flush_to(nr);
    }
}

// one warp per CHUNK of a long row -> partial[chunk, :]
template <int NC>
__global__ void __launch_bounds__(256)
spmm_chunk_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                  const double* __restrict__ values, const double* __restrict__ X,
                  const int32_t* __restrict__ chunk_row, const int32_t* __restrict__ chunk_start, int64_t nchunks,
                  int k, int c0, double* __restrict__ partial) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * 256) >> 5;
    for (int64_t ch = warp0; ch < nchunks; ch += nwarp) {
        const int a = chunk_start[ch], b = min(a + CHUNK, rowptr[chunk_row[ch] + 1]);
        double acc[NC];
#pragma unroll
        for (int q = 0; q < NC; ++q) acc[q] = 0.0;
        spmm_range<NC>(colind, values, X, k, c0, a, b, lane, acc);
#pragma unroll
        for (int u = 0; u < NC; ++u) {
            const int cc = lane + 32 * u;
            if (c0 + cc < k) partial[(size_t)ch * k + c0 + cc] = acc[u];
        }
    }
}
// one warp per long row: chunk partials added in chunk order (deterministic)
__global__ void __launch_bounds__(256)
spmm_combine_kernel(const int32_t* __restrict__ long_rows, const int32_t* __restrict__ chunk0, int64_t nlong, int k,
                    const double* __restrict__ partial, double* __restrict__ Y) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    if (w >= nlong) return;
    const int c_a = chunk0[w], c_b = chunk0[w + 1];
    for (int cc = lane; cc < k; cc += 32) {
        double s = 0.0;
        for (int ch = c_a; ch < c_b; ++ch) s += partial[(size_t)ch * k + cc];
        Y[(size_t)long_rows[w] * k + cc] = s;
    }
}

int launch_spmm(const Csr* c_, const double* X, int64_t k, double* Y) {
    Csr* c = const_cast<Csr*>(c_);
    if (c->m == 0 || k == 0) return AB_OK;
    if (k == 1) return launch_spmv(c, X, Y, nullptr, nullptr, nullptr, nullptr, nullptr);
    if (k > INT32_MAX / 2) return fail(AB_EUNSUPPORTED, "k too large");
    const int kk = (int)k;
    if (kk >= 32) {
        AB_TRY(csr_ensure_long(c));
        const int thresh = c->long_nrows > 0 ? LONG_ROW : INT32_MAX;
        int64_t nblk = (c->m + 7) / 8;
        int64_t cap  = (int64_t)sm_count() * 8 * 8;
        unsigned grid = (unsigned)(nblk < cap ? nblk : cap);
        DevBuf<double> partial;
        if (c->long_nrows > 0) AB_TRY(partial.alloc((size_t)c->long_nchunks * kk));
        unsigned gridc = (unsigned)((c->long_nchunks + 7) / 8);
        // rows of a few entries dominate (average <= 24 per row): a warp per group of 8 rows keeps the gathers in flight
        const bool grouped = opt("spmm_group", 1.0) != 0.0 && (double)c->nnz <= 24.0 * (double)c->m;
        int64_t nblkg = ((c->m + SPMM_GROUP - 1) / SPMM_GROUP + 7) / 8;
        unsigned gridg = (unsigned)(nblkg < cap ? (nblkg > 0 ? nblkg : 1) : cap);
        for (int c0 = 0; c0 < kk; c0 += 128) {
            const int rem = kk - c0;
#define AB_W(NCC)                                                                                                   \
    do {                                                                                                            \
        if (grouped)                                                                                                \
            spmm_group_kernel<NCC><<<gridg, 256, 0, stream()>>>(c->rowptr, c->colind, c->values, X, Y, c->m, kk, c0, thresh); \
        else                                                                                                        \
            spmm_warp_kernel<NCC><<<grid, 256, 0, stream()>>>(c->rowptr, c->colind, c->values, X, Y, c->m, kk, c0, thresh); \
        AB_LAUNCH_CHECK();                                                                                          \
        if (c->long_nrows > 0) {                                                                                    \
            spmm_chunk_kernel<NCC><<<gridc, 256, 0, stream()>>>(c->rowptr, c->colind, c->values, X, c->long_chunk_row, \
                                                                c->long_chunk_start, c->long_nchunks, kk, c0, partial.p); \
            AB_LAUNCH_CHECK();                                                                                      \
        }                                                                                                           \
    } while (0)
            if (rem > 96) AB_W(4);
            else if (rem > 64) AB_W(3);
            else if (rem > 32) AB_W(2);
            else AB_W(1);
#undef AB_W
        }
        if (c->long_nrows > 0) {
            spmm_combine_kernel<<<(unsigned)((c->long_nrows + 7) / 8), 256, 0, stream()>>>(c->long_rows, c->long_row_chunk0,
                                                                                         c->long_nrows, kk, partial.p, Y);
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
#define AB_SPMM(CLL) spmm_kernel<CLL><<<grid, 256, 0, stream()>>>(c->rowptr, c->colind, c->values, X, Y, c->m, kk)
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

// ---------------------------------------------------------------- SDDMM (value gradient)
// small k: CL lanes per stored entry, dv[e] = <dY[row_e,:], X[col_e,:]>
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

// 32 <= k <= 128: a warp handles entries [a,b) of one row.  The dY row stays in registers; 32 entries
// at once: lane c accumulates its columns' products for each entry (32 independent X-row gathers in
// flight), then a butterfly transpose-reduce (31 shuffles per 32 entries instead of 160) leaves
// entry l's dot product in lane l, stored coalesced.
template <int NC>
__device__ __forceinline__ void sddmm_range(const int32_t* __restrict__ colind, const double* __restrict__ X, int k,
                                            const double (&dy)[NC], int a, int b, int lane, double* __restrict__ dv) {
    constexpr int EB = 16;   // entries per batch
    for (int base = a; base < b; base += EB) {
        const int cnt = min(EB, b - base);
        // lanes beyond the batch keep column 0: their gathers hit row 0 of X (cached) and their
        // products are never stored, so every load below is unconditional and all EB of them are
        // in flight together (a predicated load inside a branch would serialise the warp)
        const int cj = lane < cnt ? colind[base + lane] : 0;
        double prod[EB];
#pragma unroll
        for (int q = 0; q < EB; ++q) {
            const int cq = __shfl_sync(0xffffffffu, cj, q);
            const double* xr = X + (size_t)cq * k;
            double s = 0.0;
#pragma unroll
            for (int u = 0; u < NC; ++u) {
                const int cc = lane + 32 * u;
                const double xv = __ldg(xr + (cc < k ? cc : 0));
                s = cc < k ? fma(dy[u], xv, s) : s;
            }
            prod[q] = s;
        }
#pragma unroll
        for (int s = EB / 2; s >= 1; s >>= 1) {
            const bool up = (lane & s) != 0;
#pragma unroll
            for (int i = 0; i < s; ++i) {
                const double send = up ? prod[i] : prod[i + s];
                const double keep = up ? prod[i + s] : prod[i];
                prod[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
            }
        }
        // lanes l and l^16 each hold half of entry (l & 15)'s sum
        const double tot = prod[0] + __shfl_xor_sync(0xffffffffu, prod[0], 16);
        if (lane < cnt) dv[base + lane] = tot;
    }
}

template <int NC>
__global__ void __launch_bounds__(256)
sddmm_warp_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                  const double* __restrict__ dY, const double* __restrict__ X, int64_t m, int k,
                  const int32_t* __restrict__ chunk_row, const int32_t* __restrict__ chunk_start, int64_t nchunks,
                  int long_thresh, double* __restrict__ dv) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const int64_t nwarp = ((int64_t)gridDim.x * 256) >> 5;
    // work items: the m rows (long ones skipped) followed by the chunks of the long rows
    for (int64_t w = warp0; w < m + nchunks; w += nwarp) {
        int64_t row; int a, b;
        if (w < m) {
            row = w; a = rowptr[row]; b = rowptr[row + 1];
            if (a == b || b - a > long_thresh) continue;
        } else {
            const int64_t ch = w - m;
            row = chunk_row[ch]; a = chunk_start[ch]; b = min(a + CHUNK, rowptr[row + 1]);
        }
        double dy[NC];
#pragma unroll
        for (int u = 0; u < NC; ++u) {
            const int cc = lane + 32 * u;
            dy[u] = cc < k ? __ldg(dY + (size_t)row * k + cc) : 0.0;
        }
        sddmm_range<NC>(colind, X, k, dy, a, b, lane, dv);
    }
}

// 32 <= k <= 128, entry-parallel form: a warp takes 16 consecutive stored ENTRIES whatever rows they belong to — on a
// power-law matrix most rows hold 1-3 entries, and a warp per row keeps one or two 256-byte X gathers in flight.  Here
// every warp has 16 X-row gathers and 16 dY-row loads in flight (consecutive entries share their dY row: L1 hits);
// long rows need no chunking.  Same per-entry arithmetic and the same butterfly as sddmm_range: identical bits.
template <int NC>
__global__ void __launch_bounds__(256)
sddmm_entry_kernel(const int32_t* __restrict__ entry_row, const int32_t* __restrict__ colind,
                   const double* __restrict__ dY, const double* __restrict__ X, size_t nnz, int k,
                   double* __restrict__ dv) {
    constexpr int EB = 16;
    const int lane = threadIdx.x & 31;
    const size_t warp0 = ((size_t)blockIdx.x * 256 + threadIdx.x) >> 5;
    const size_t nwarp = ((size_t)gridDim.x * 256) >> 5;
    for (size_t base = warp0 * EB; base < nnz; base += nwarp * EB) {
        const int cnt = (int)min((size_t)EB, nnz - base);
        const int cj = lane < cnt ? colind[base + lane] : 0;
        const int rj = lane < cnt ? entry_row[base + lane] : 0;
        double prod[EB];
#pragma unroll
        for (int q = 0; q < EB; ++q) {
            const int cq = __shfl_sync(0xffffffffu, cj, q), rq = __shfl_sync(0xffffffffu, rj, q);
            const double* xr = X + (size_t)cq * k;
            const double* yr = dY + (size_t)rq * k;
            double s = 0.0;
#pragma unroll
            for (int u = 0; u < NC; ++u) {
                const int cc = lane + 32 * u;
                const double xv = __ldg(xr + (cc < k ? cc : 0));
                const double yv = __ldg(yr + (cc < k ? cc : 0));
                s = cc < k ? fma(yv, xv, s) : s;
            }
            prod[q] = s;
        }
#pragma unroll
        for (int s = EB / 2; s >= 1; s >>= 1) {
            const bool up = (lane & s) != 0;
#pragma unroll
            for (int i = 0; i < s; ++i) {
                const double send = up ? prod[i] : prod[i + s];
                const double keep = up ? prod[i + s] : prod[i];
                prod[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
            }
        }
        const double tot = prod[0] + __shfl_xor_sync(0xffffffffu, prod[0], 16);
        if (lane < cnt) dv[base + lane] = tot;
    }
}

__global__ void gather_slot_kernel(const int32_t* __restrict__ slot, const double* __restrict__ dv, size_t T,
                                   double* __restrict__ out) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < T) out[t] = dv[slot[t]];
}

static int launch_sddmm(Csr* c, const double* dY, const double* X, int k, double* dv) {
    if (c->nnz == 0) return AB_OK;
    if (k >= 32 && k <= 128 && opt("sddmm_entry", 1.0) != 0.0) {
        AB_TRY(csr_ensure_entry_row(c));
        const size_t nwork = ((size_t)c->nnz + 15) / 16;
        size_t nblk = (nwork + 7) / 8, cap = (size_t)sm_count() * 8 * 8;
        const unsigned grid = (unsigned)(nblk < cap ? nblk : cap);
#define AB_SDE(NCC) sddmm_entry_kernel<NCC><<<grid, 256, 0, stream()>>>(c->entry_row, c->colind, dY, X, (size_t)c->nnz, k, dv)
        if (k <= 32) AB_SDE(1);
        else if (k <= 64) AB_SDE(2);
        else if (k <= 96) AB_SDE(3);
        else AB_SDE(4);
#undef AB_SDE
        AB_LAUNCH_CHECK();
        return AB_OK;
    }
    if (k >= 32 && k <= 128) {
        AB_TRY(csr_ensure_long(c));
        const int thresh = c->long_nrows > 0 ? LONG_ROW : INT32_MAX;
        int64_t nblk = (c->m + c->long_nchunks + 7) / 8;
        int64_t cap  = (int64_t)sm_count() * 8 * 8;
        unsigned grid = (unsigned)(nblk < cap ? nblk : cap);
#define AB_SDW(NCC)                                                                                                \
    sddmm_warp_kernel<NCC><<<grid, 256, 0, stream()>>>(c->rowptr, c->colind, dY, X, c->m, k, c->long_chunk_row,    \
                                                       c->long_chunk_start, c->long_nchunks, thresh, dv)
        if (k <= 32) AB_SDW(1);
        else if (k <= 64) AB_SDW(2);
        else if (k <= 96) AB_SDW(3);
        else AB_SDW(4);
#undef AB_SDW
        AB_LAUNCH_CHECK();
        return AB_OK;
    }
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
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const bool timing = opt("timers", 0.0) != 0.0;
    if (timing) { AB_CUDA(cudaEventCreate(&e0)); AB_CUDA(cudaEventCreate(&e1)); AB_CUDA(cudaEventRecord(e0, stream())); }
    Csr* a = c;
    if (transA) AB_TRY(csr_ensure_transpose(c, &a));
    if ((a->n * k > 0 && !X_) || (a->m * k > 0 && !Y_)) return fail(AB_EINVAL, "null operand");
    In<double> X; Out<double> Y;
    AB_TRY(X.bind(X_, (size_t)(a->n * k))); AB_TRY(Y.bind(Y_, (size_t)(a->m * k)));
    AB_TRY(launch_spmm(a, X.d, k, Y.d));
    AB_TRY(Y.commit());
    if (X.tmp.p) AB_CUDA(cudaStreamSynchronize(stream()));
    if (timing) {
        AB_CUDA(cudaEventRecord(e1, stream())); AB_CUDA(cudaEventSynchronize(e1));
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        timer_add(AB_TIMER_SPMM, ms * 1e-3);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    }
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
    AB_TRY(csr_ensure_slot(c));
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
    AB_TRY(csr_ensure_work(c, 10));
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
