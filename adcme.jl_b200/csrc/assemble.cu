// assemble.cu — COO -> CSR assembly with duplicate summation (segmented sort-and-reduce),
// `compress`, mpi_create_matrix's COO -> (rows,ncols,cols), and the SparseAssembler.
//
// Reference behaviour being reproduced (see oracle/adcme_oracle.c for the CPU restatement):
//   - Eigen setFromTriplets as called at deps/CustomOps/SparseSolver/SparseSolver.h:16-22:
//     duplicates summed in INPUT order, inner indices sorted;
//   - SparseCompressor, deps/CustomOps/SparseCompress/SparseCompress.h:21-54;
//   - MPICreateMatrix_forward, deps/Plugin/MPITensor/MPITensor.h:1-36;
//   - SparseAccum, deps/CustomOps/SparseAccumulate/Impl.cpp:11-48.
//
// Device algorithm: pack (row,col) into one 64-bit key, STABLE least-significant-digit radix
// sort (8-bit digits, only over the bits the matrix shape needs) carrying the triplet index,
// head-flag scan to name the merged entries, then one thread per stored entry adds its
// duplicates left to right — i.e. in input order, which makes the fp64 sums bit-identical to
// the reference's sequential accumulation.
#include "common.cuh"
#include <map>

namespace ab {

// =====================================================================================
// exclusive scan (uint32)
// =====================================================================================
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS   = 8;
constexpr int SCAN_CHUNK   = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS)
scan_block_sums_kernel(const uint32_t* __restrict__ in, size_t n, uint32_t* __restrict__ sums) {
    __shared__ uint32_t sh[SCAN_THREADS / 32];
    size_t base = (size_t)blockIdx.x * SCAN_CHUNK;
    uint32_t a  = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        size_t idx = base + (size_t)i * SCAN_THREADS + threadIdx.x;
        if (idx < n) a += in[idx];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t v = threadIdx.x < SCAN_THREADS / 32 ? sh[threadIdx.x] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) sums[blockIdx.x] = v;
    }
}

// scans one chunk; adds offsets[blockIdx.x] when given; writes grand total when asked (1 block)
__global__ void __launch_bounds__(SCAN_THREADS)
scan_chunk_kernel(const uint32_t* in, uint32_t* out, size_t n, const uint32_t* __restrict__ offsets,
                  uint32_t* total_out) {
    __shared__ uint32_t buf[SCAN_CHUNK];
    __shared__ uint32_t wsum[SCAN_THREADS / 32];
    const size_t base = (size_t)blockIdx.x * SCAN_CHUNK;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        size_t idx = base + (size_t)i * SCAN_THREADS + threadIdx.x;
        buf[i * SCAN_THREADS + threadIdx.x] = idx < n ? in[idx] : 0u;
    }
    __syncthreads();
    uint32_t loc[SCAN_ITEMS];
    uint32_t tsum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        loc[i] = buf[threadIdx.x * SCAN_ITEMS + i];
        tsum += loc[i];
    }
    // inclusive warp scan of tsum
    uint32_t inc = tsum;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    if (w == 0) {
        uint32_t v = lane < SCAN_THREADS / 32 ? wsum[lane] : 0;
        uint32_t vi = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, vi, o);
            if (lane >= o) vi += t;
        }
        if (lane < SCAN_THREADS / 32) wsum[lane] = vi - v;  // exclusive warp offsets
        if (lane == SCAN_THREADS / 32 - 1 && total_out) *total_out = vi + (offsets ? offsets[blockIdx.x] : 0u);
    }
    __syncthreads();
    uint32_t run = wsum[w] + (inc - tsum) + (offsets ? offsets[blockIdx.x] : 0u);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        buf[threadIdx.x * SCAN_ITEMS + i] = run;
        run += loc[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        size_t idx = base + (size_t)i * SCAN_THREADS + threadIdx.x;
        if (idx < n) out[idx] = buf[i * SCAN_THREADS + threadIdx.x];
    }
}

int scan_exclusive_u32(const uint32_t* in, uint32_t* out, size_t n, uint32_t* d_total) {
    if (n == 0) {
        if (d_total) AB_CUDA(cudaMemsetAsync(d_total, 0, sizeof(uint32_t), stream()));
        return AB_OK;
    }
    size_t nb = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
    if (nb == 1) {
        scan_chunk_kernel<<<1, SCAN_THREADS, 0, stream()>>>(in, out, n, nullptr, d_total);
        AB_LAUNCH_CHECK();
        return AB_OK;
    }
    DevBuf<uint32_t> sums;
    AB_TRY(sums.alloc(nb));
    scan_block_sums_kernel<<<(unsigned)nb, SCAN_THREADS, 0, stream()>>>(in, n, sums.p);
    AB_LAUNCH_CHECK();
    AB_TRY(scan_exclusive_u32(sums.p, sums.p, nb, d_total));
    scan_chunk_kernel<<<(unsigned)nb, SCAN_THREADS, 0, stream()>>>(in, out, n, sums.p, nullptr);
    AB_LAUNCH_CHECK();
    return AB_OK;
}

// =====================================================================================
// stable LSD radix sort, 64-bit keys + 32-bit payload
// =====================================================================================
constexpr int RS_THREADS = 256;
constexpr int RS_WARPS   = RS_THREADS / 32;
constexpr int RS_ITEMS   = 16;
constexpr int RS_TILE    = RS_THREADS * RS_ITEMS;   // 4096 keys per CTA
constexpr int RS_RADIX   = 256;
constexpr int RS_SMEM_BYTES = RS_TILE * 12 + (RS_WARPS + 2) * RS_RADIX * 4 + RS_WARPS * 4 + 16;

__global__ void __launch_bounds__(RS_THREADS)
radix_hist_kernel(const uint64_t* __restrict__ keys, size_t n, int shift, uint32_t* __restrict__ counts,
                  unsigned ntiles) {
    __shared__ uint32_t hist[RS_RADIX];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const size_t base = (size_t)blockIdx.x * RS_TILE;
    uint64_t k[RS_ITEMS];
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {          // all loads in flight before the first atomic
        size_t idx = base + (size_t)i * RS_THREADS + threadIdx.x;
        k[i] = idx < n ? keys[idx] : 0ull;
    }
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        size_t idx = base + (size_t)i * RS_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&hist[(unsigned)(k[i] >> shift) & (RS_RADIX - 1)], 1u);
    }
    __syncthreads();
    counts[(size_t)threadIdx.x * ntiles + blockIdx.x] = hist[threadIdx.x];
}

__global__ void __launch_bounds__(RS_THREADS, 3)
radix_scatter_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, size_t n,
                     int shift, const uint32_t* __restrict__ offsets, unsigned ntiles) {
    // dynamic shared memory (RS_SMEM_BYTES = 58 KB > the 48 KB static limit)
    extern __shared__ __align__(16) unsigned char rs_smem[];
    uint64_t* skey = reinterpret_cast<uint64_t*>(rs_smem);                       // [RS_TILE]
    uint32_t* sval = reinterpret_cast<uint32_t*>(skey + RS_TILE);                // [RS_TILE]
    uint32_t (*whist)[RS_RADIX] = reinterpret_cast<uint32_t (*)[RS_RADIX]>(sval + RS_TILE);   // [RS_WARPS][RS_RADIX]
    uint32_t* dbase = &whist[0][0] + RS_WARPS * RS_RADIX;                        // [RS_RADIX]
    uint32_t* gbase = dbase + RS_RADIX;                                          // [RS_RADIX]
    uint32_t* wtot  = gbase + RS_RADIX;                                          // [RS_WARPS]
    for (int i = threadIdx.x; i < RS_WARPS * RS_RADIX; i += RS_THREADS) (&whist[0][0])[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const size_t wbase = (size_t)blockIdx.x * RS_TILE + (size_t)w * (32 * RS_ITEMS);
    uint64_t key[RS_ITEMS];
    uint32_t rank[RS_ITEMS], val[RS_ITEMS];
    // items of a warp are visited in global index order -> ranks are stable
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        size_t idx = wbase + (size_t)i * 32 + lane;
        bool valid = idx < n;
        key[i]     = valid ? keys_in[idx] : ~0ull;
        val[i]     = (valid && vals_in) ? vals_in[idx] : (uint32_t)idx;
    }
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        size_t idx = wbase + (size_t)i * 32 + lane;
        bool valid = idx < n;
        unsigned act = __ballot_sync(0xffffffffu, valid);
        unsigned d = (unsigned)(key[i] >> shift) & (RS_RADIX - 1);
        unsigned peers = 0, prev = 0;
        if (valid) {
            peers = __match_any_sync(act, d);
            prev  = whist[w][d];
        }
        __syncwarp();
        if (valid) {
            rank[i] = prev + __popc(peers & lt);
            if (lane == __ffs(peers) - 1) whist[w][d] = prev + __popc(peers);
        }
        __syncwarp();
    }
    __syncthreads();
    {   // digit d = threadIdx.x: per-warp counts -> bases inside the digit; digit totals -> tile-local
        // exclusive offsets (block scan); gbase[d] + (position in the locally sorted tile) = global slot
        const unsigned d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ++ww) {
            uint32_t t   = whist[ww][d];
            whist[ww][d] = run;
            run += t;
        }
        uint32_t inc = run;                       // inclusive scan of the 256 digit totals
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wtot[w] = inc;
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ++ww) woff += (ww < w) ? wtot[ww] : 0u;
        const uint32_t excl = woff + inc - run;
        dbase[d] = excl;
        gbase[d] = offsets[(size_t)d * ntiles + blockIdx.x] - excl;
    }
    __syncthreads();
    // local sort: every key goes to its slot of the digit-ordered tile in shared memory
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        size_t idx = wbase + (size_t)i * 32 + lane;
        if (idx < n) {
            unsigned d    = (unsigned)(key[i] >> shift) & (RS_RADIX - 1);
            uint32_t tpos = dbase[d] + whist[w][d] + rank[i];
            skey[tpos] = key[i];
            sval[tpos] = val[i];
        }
    }
    __syncthreads();
    // coalesced write-out: consecutive threads hold consecutive slots of the same digit run
    const size_t tbase = (size_t)blockIdx.x * RS_TILE;
    const uint32_t nvalid = (uint32_t)min((size_t)RS_TILE, n - tbase);
#pragma unroll
    for (int i = 0; i < RS_ITEMS; ++i) {
        const uint32_t li = (uint32_t)i * RS_THREADS + threadIdx.x;
        if (li < nvalid) {
            const uint64_t kk = skey[li];
            const unsigned d  = (unsigned)(kk >> shift) & (RS_RADIX - 1);
            const uint32_t pos = gbase[d] + li;
            keys_out[pos] = kk;
            vals_out[pos] = sval[li];
        }
    }
}

int radix_sort_u64(uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp, uint32_t* vals_tmp, size_t n,
                   int nbits, uint64_t** keys_out, uint32_t** vals_out, int bit0) {
    // vals == identity on entry is implied: the first pass synthesises the payload from the index.
    if (n >= (size_t)UINT32_MAX) return fail(AB_EUNSUPPORTED, "radix sort: %zu elements exceed 32-bit positions", n);
    uint64_t* kin = keys; uint64_t* kout = keys_tmp;
    uint32_t* vin = nullptr; uint32_t* vout = vals_tmp;   // nullptr payload -> identity
    uint32_t* vother = vals;
    if (n == 0) { *keys_out = keys; *vals_out = vals; return AB_OK; }
    unsigned ntiles = (unsigned)((n + RS_TILE - 1) / RS_TILE);
    DevBuf<uint32_t> counts;
    AB_TRY(counts.alloc((size_t)RS_RADIX * ntiles));
    int npass = (nbits + 7) / 8;
    if (npass < 1) npass = 1;
    static bool smem_attr_set = false;
    if (!smem_attr_set) {
        AB_CUDA(cudaFuncSetAttribute(radix_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_SMEM_BYTES));
        smem_attr_set = true;
    }
    for (int p = 0; p < npass; ++p) {
        int shift = bit0 + 8 * p;
        radix_hist_kernel<<<ntiles, RS_THREADS, 0, stream()>>>(kin, n, shift, counts.p, ntiles);
        AB_LAUNCH_CHECK();
        AB_TRY(scan_exclusive_u32(counts.p, counts.p, (size_t)RS_RADIX * ntiles, nullptr));
        radix_scatter_kernel<<<ntiles, RS_THREADS, RS_SMEM_BYTES, stream()>>>(kin, vin, kout, vout, n, shift, counts.p, ntiles);
        AB_LAUNCH_CHECK();
        // rotate buffers
        uint64_t* tk = kin; kin = kout; kout = tk;
        uint32_t* newin = vout;
        vout = (vin == nullptr) ? vother : vin;
        vin  = newin;
    }
    *keys_out = kin;
    *vals_out = vin;
    return AB_OK;
}

// =====================================================================================
// COO -> CSR
// =====================================================================================
static int bits_for(int64_t dim) {   // bits needed to store values in [0, dim)
    int b = 1;
    while (b < 63 && ((int64_t)1 << b) < dim) ++b;
    return b;
}

template <class IdxT>
__global__ void make_keys_kernel(const IdxT* __restrict__ ii, const IdxT* __restrict__ jj, size_t T,
                                 int64_t m, int64_t n, int base, int colbits, uint64_t* __restrict__ keys,
                                 int* __restrict__ err) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    int64_t r = (int64_t)ii[t] - base, c = (int64_t)jj[t] - base;
    if (r < 0 || r >= m || c < 0 || c >= n) {
        atomicExch(err, 1);
        r = 0; c = 0;
    }
    keys[t] = ((uint64_t)r << colbits) | (uint64_t)c;
}

__global__ void head_flags_kernel(const uint64_t* __restrict__ keys, size_t T, uint32_t* __restrict__ flags) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= T) return;
    flags[k] = (k == 0 || keys[k] != keys[k - 1]) ? 1u : 0u;
}

// excl[k] = number of heads before k.  Head k owns stored entry e = excl[k].
__global__ void emit_entries_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ excl,
                                    size_t T, int64_t m, int colbits,
                                    int32_t* __restrict__ colind, uint32_t* __restrict__ segstart,
                                    int32_t* __restrict__ rowptr, int32_t* __restrict__ entry_row) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= T) return;
    const uint64_t key  = keys[k];
    const uint64_t prev = k ? keys[k - 1] : 0ull;
    const bool head     = (k == 0) || (key != prev);
    const uint32_t e    = head ? excl[k] : excl[k] - 1;
    const int64_t r     = (int64_t)(key >> colbits);
    if (head) {
        colind[e]    = (int32_t)(key & ((1ull << colbits) - 1));
        segstart[e]  = (uint32_t)k;
        entry_row[e] = (int32_t)r;
        // entries are sorted by row: this head opens every row after the previous element's row
        const int64_t rp = k ? (int64_t)(prev >> colbits) : -1;
        for (int64_t rr = rp + 1; rr <= r; ++rr) rowptr[rr] = (int32_t)e;
    }
    if (k == T - 1)
        for (int64_t rr = r + 1; rr <= m; ++rr) rowptr[rr] = (int32_t)(e + 1);   // rows after the last entry
}

// slot[t] = stored entry of input triplet t (lazy: only gradient delivery and ab_csr_export_map need it)
__global__ void slot_kernel(const uint32_t* __restrict__ segstart, const uint32_t* __restrict__ perm, size_t nnz,
                            int32_t* __restrict__ slot) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    for (uint32_t k = segstart[e]; k < segstart[e + 1]; ++k) slot[perm[k]] = (int32_t)e;
}

// one thread per stored entry: left-to-right (= input order) sum of its duplicates
__global__ void reduce_segments_kernel(const uint32_t* __restrict__ segstart, const uint32_t* __restrict__ perm,
                                       const double* __restrict__ vv, size_t nnz, double* __restrict__ values) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    uint32_t a = segstart[e], b = segstart[e + 1];
    double s = vv[perm[a]];
    for (uint32_t k = a + 1; k < b; ++k) s += vv[perm[k]];
    values[e] = s;
}

__global__ void tile_facts_kernel(const int32_t* __restrict__ rowptr, int64_t m, int* __restrict__ facts) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int len = 0;
    if (r < m) len = rowptr[r + 1] - rowptr[r];
    // one block = 256 rows = one / two / four SpMV tiles of 256 / 128 / 64 rows
    __shared__ int smax[8];
    int v = len;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        int mx = 0;
        for (int i = 0; i < 8; ++i) mx = max(mx, smax[i]);
        atomicMax(&facts[3], mx);
        const int64_t r0 = (int64_t)blockIdx.x * 256;
        for (int lvl = 0; lvl < 3; ++lvl) {
            const int R = 256 >> lvl;
            int best = 0;
            for (int64_t a = r0; a < min(r0 + 256, m); a += R) {
                int64_t b = min(a + R, m);
                best = max(best, rowptr[b] - rowptr[a]);
            }
            atomicMax(&facts[lvl], best);
        }
    }
}

int csr_finalize(Csr* c) {
    c->dinv_valid = false;
    if (c->m == 0) { c->max_tile_nnz[0] = c->max_tile_nnz[1] = c->max_tile_nnz[2] = 0; c->max_row_nnz = 0; return AB_OK; }
    DevBuf<int> facts;
    AB_TRY(facts.alloc(4));
    AB_CUDA(cudaMemsetAsync(facts.p, 0, 4 * sizeof(int), stream()));
    unsigned nb = (unsigned)((c->m + 255) / 256);
    tile_facts_kernel<<<nb, 256, 0, stream()>>>(c->rowptr, c->m, facts.p);
    AB_LAUNCH_CHECK();
    int h[4];
    AB_CUDA(cudaMemcpyAsync(h, facts.p, sizeof(h), cudaMemcpyDeviceToHost, stream()));
    AB_CUDA(cudaStreamSynchronize(stream()));
    c->max_tile_nnz[0] = h[0]; c->max_tile_nnz[1] = h[1]; c->max_tile_nnz[2] = h[2];
    c->max_row_nnz  = h[3];
    return AB_OK;
}

// ---- rows first, then columns inside each row ------------------------------------------------------------------------
// The keys are (row << colbits) | col.  Sorting all rowbits + colbits bits costs six 8-bit passes at 256^3 (48 bits), each
// moving 32 bytes per triplet.  Sorting by the ROW bits only (three passes; stable, so every row keeps its triplets in
// input order) and then ordering the few entries of each row by column in place gives the same array bit for bit: one
// thread per row walks its short segment with a stable insertion sort (ties keep input order = what the LSD passes give).
// Rows with more than ROWSORT_CAP triplets (power-law matrices) raise `overflow` and the caller falls back to the full sort.
constexpr int ROWSORT_CAP = 48;
__global__ void row_starts_kernel(const uint64_t* __restrict__ keys, size_t T, int64_t m, int colbits, uint32_t* __restrict__ rowstart) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= T) return;
    const int64_t r = (int64_t)(keys[k] >> colbits), rp = k ? (int64_t)(keys[k - 1] >> colbits) : -1;
    for (int64_t rr = rp + 1; rr <= r; ++rr) rowstart[rr] = (uint32_t)k;     // rows (rp, r] start here (all but r are empty)
    if (k == T - 1) for (int64_t rr = r + 1; rr <= m; ++rr) rowstart[rr] = (uint32_t)T;
}
// One CTA = 128 consecutive rows = one contiguous piece of the row-sorted arrays: staged into shared memory by coalesced
// loads, every thread orders its own row there, coalesced write-back.  (A first version that walked the rows straight in
// global / local memory took 6.8 ms at C2 — as much as the three saved passes.)  A CTA whose piece exceeds the stage
// (ROWSORT_STAGE entries) sorts its rows in place through local arrays instead.
constexpr int ROWSORT_ROWS = 128;
constexpr int ROWSORT_STAGE = 3072;
__device__ __forceinline__ void rowsort_insertion(uint64_t* kk, uint32_t* vv, int len) {
    for (int i = 1; i < len; ++i) {                 // stable: an element moves left only past STRICTLY larger keys
        const uint64_t ki = kk[i]; const uint32_t vi = vv[i];
        int j = i - 1;
        while (j >= 0 && kk[j] > ki) { kk[j + 1] = kk[j]; vv[j + 1] = vv[j]; --j; }
        kk[j + 1] = ki; vv[j + 1] = vi;
    }
}
__global__ void __launch_bounds__(ROWSORT_ROWS)
row_colsort_kernel(uint64_t* __restrict__ keys, uint32_t* __restrict__ idx, const uint32_t* __restrict__ rowstart, int64_t m,
                   int* __restrict__ overflow) {
    __shared__ uint64_t sk[ROWSORT_STAGE];
    __shared__ uint32_t sv[ROWSORT_STAGE];
    __shared__ int s_changed;
    const int t = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * ROWSORT_ROWS, r1 = min(r0 + ROWSORT_ROWS, m);
    const uint32_t a0 = rowstart[r0], b0 = rowstart[r1];
    const uint32_t n = b0 - a0;
    const int64_t r = r0 + t;
    uint32_t a = 0, b = 0;
    if (r < r1) { a = rowstart[r]; b = rowstart[r + 1]; }
    const int len = (int)(b - a);
    if (len > ROWSORT_CAP) atomicExch(overflow, 1);
    if (n <= (uint32_t)ROWSORT_STAGE) {
        if (t == 0) s_changed = 0;
        for (uint32_t i = t; i < n; i += ROWSORT_ROWS) { sk[i] = keys[a0 + i]; sv[i] = idx[a0 + i]; }
        __syncthreads();
        if (len >= 2 && len <= ROWSORT_CAP) {
            uint64_t* kk = sk + (a - a0);
            uint32_t* vv = sv + (a - a0);
            bool sorted = true;
            for (int i = 1; i < len; ++i) sorted = sorted && kk[i - 1] <= kk[i];
            if (!sorted) { rowsort_insertion(kk, vv, len); s_changed = 1; }
        }
        __syncthreads();
        if (s_changed)
            for (uint32_t i = t; i < n; i += ROWSORT_ROWS) { keys[a0 + i] = sk[i]; idx[a0 + i] = sv[i]; }
        return;
    }
    if (len < 2 || len > ROWSORT_CAP) return;
    uint64_t kk[ROWSORT_CAP];
    uint32_t vv[ROWSORT_CAP];
    bool sorted = true;
    for (int i = 0; i < len; ++i) {
        kk[i] = keys[a + i]; vv[i] = idx[a + i];
        sorted = sorted && (i == 0 || kk[i - 1] <= kk[i]);
    }
    if (sorted) return;
    rowsort_insertion(kk, vv, len);
    for (int i = 0; i < len; ++i) { keys[a + i] = kk[i]; idx[a + i] = vv[i]; }
}

template <class IdxT>
static int csr_from_coo_impl(int64_t T, const IdxT* ii_, const IdxT* jj_, const double* vv_, int64_t m,
                             int64_t n, int base, int64_t* out_h) {
    AB_TRY(ensure_device());
    if (!out_h) return fail(AB_EINVAL, "out_h is null");
    if (T < 0 || m < 0 || n < 0) return fail(AB_EINVAL, "negative size");
    if (T > 0 && (!ii_ || !jj_ || !vv_)) return fail(AB_EINVAL, "null triplet array");
    if (m >= INT32_MAX || n >= INT32_MAX) return fail(AB_EUNSUPPORTED, "matrix dimension exceeds int32 local index space");
    if (T >= (int64_t)UINT32_MAX) return fail(AB_EUNSUPPORTED, "more than 2^32-1 triplets in one call");
    std::lock_guard<std::mutex> lk(big_lock());
    cudaEvent_t e0, e1;
    AB_CUDA(cudaEventCreate(&e0)); AB_CUDA(cudaEventCreate(&e1));
    AB_CUDA(cudaEventRecord(e0, stream()));

    In<IdxT> ii, jj; In<double> vv;
    AB_TRY(ii.bind(ii_, T)); AB_TRY(jj.bind(jj_, T)); AB_TRY(vv.bind(vv_, T));

    Csr* c = new Csr();
    c->m = m; c->n = n; c->T = T;
    int rc = AB_OK;
    auto body = [&]() -> int {
        const int colbits = bits_for(n > 0 ? n : 1), rowbits = bits_for(m > 0 ? m : 1);
        DevBuf<uint64_t> k0, k1;
        DevBuf<uint32_t> p0, p1, flags, rowcnt, total;
        DevBuf<int> err;
        AB_TRY(k0.alloc(T)); AB_TRY(k1.alloc(T)); AB_TRY(p0.alloc(T)); AB_TRY(p1.alloc(T));
        AB_TRY(err.alloc(1)); AB_TRY(total.alloc(1));
        AB_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), stream()));
        uint64_t* ks = k0.p; uint32_t* ps = p0.p;
        uint32_t nnz = 0;
        if (T > 0) {
            unsigned nb = (unsigned)((T + 255) / 256);
            make_keys_kernel<IdxT><<<nb, 256, 0, stream()>>>(ii.d, jj.d, (size_t)T, m, n, base, colbits, k0.p, err.p);
            AB_LAUNCH_CHECK();
            bool done_sort = false;
            // rows first, columns inside each row (see row_colsort_kernel): half the passes of the full-key sort
            if (opt("asm_rowsort", 1.0) != 0.0 && colbits >= 8 && T >= 65536 && (double)T <= 24.0 * (double)(m > 0 ? m : 1)) {
                DevBuf<uint32_t> rowstart; DevBuf<int> ovf;
                AB_TRY(rowstart.alloc((size_t)m + 1)); AB_TRY(ovf.alloc(1));
                AB_CUDA(cudaMemsetAsync(ovf.p, 0, sizeof(int), stream()));
                AB_TRY(radix_sort_u64(k0.p, p0.p, k1.p, p1.p, (size_t)T, rowbits, &ks, &ps, colbits));
                row_starts_kernel<<<nb, 256, 0, stream()>>>(ks, (size_t)T, m, colbits, rowstart.p);
                AB_LAUNCH_CHECK();
                row_colsort_kernel<<<(unsigned)((m + ROWSORT_ROWS - 1) / ROWSORT_ROWS), ROWSORT_ROWS, 0, stream()>>>(ks, ps, rowstart.p, m, ovf.p);
                AB_LAUNCH_CHECK();
                int hovf = 0;
                AB_CUDA(cudaMemcpyAsync(&hovf, ovf.p, sizeof(int), cudaMemcpyDeviceToHost, stream()));
                AB_CUDA(cudaStreamSynchronize(stream()));
                done_sort = hovf == 0;
                if (!done_sort) {          // a long row: start over with the full-key sort
                    make_keys_kernel<IdxT><<<nb, 256, 0, stream()>>>(ii.d, jj.d, (size_t)T, m, n, base, colbits, k0.p, err.p);
                    AB_LAUNCH_CHECK();
                }
            }
            if (!done_sort) AB_TRY(radix_sort_u64(k0.p, p0.p, k1.p, p1.p, (size_t)T, colbits + rowbits, &ks, &ps));
            AB_TRY(flags.alloc(T));
            head_flags_kernel<<<nb, 256, 0, stream()>>>(ks, (size_t)T, flags.p);
            AB_LAUNCH_CHECK();
            AB_TRY(scan_exclusive_u32(flags.p, flags.p, (size_t)T, total.p));
            int herr = 0;
            AB_CUDA(cudaMemcpyAsync(&herr, err.p, sizeof(int), cudaMemcpyDeviceToHost, stream()));
            AB_CUDA(cudaMemcpyAsync(&nnz, total.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream()));
            AB_CUDA(cudaStreamSynchronize(stream()));
            if (herr) return fail(AB_ERANGE, "triplet index outside [%d,%lld]x[%d,%lld]", base,
                                  (long long)(m + base - 1), base, (long long)(n + base - 1));
            if (nnz >= (uint32_t)INT32_MAX) return fail(AB_EUNSUPPORTED, "nnz exceeds int32");
        }
        c->nnz = nnz;
        AB_TRY(dev_alloc((void**)&c->rowptr, (size_t)(m + 1) * sizeof(int32_t)));
        AB_TRY(dev_alloc((void**)&c->colind, ((size_t)nnz + CSR_PAD) * sizeof(int32_t)));
        AB_TRY(dev_alloc((void**)&c->values, ((size_t)nnz + CSR_PAD) * sizeof(double)));
        AB_TRY(dev_alloc((void**)&c->segstart, ((size_t)nnz + 1) * sizeof(uint32_t)));
        AB_TRY(dev_alloc((void**)&c->entry_row, (size_t)(nnz > 0 ? nnz : 1) * sizeof(int32_t)));
        AB_CUDA(cudaMemsetAsync(c->colind + nnz, 0, CSR_PAD * sizeof(int32_t), stream()));
        AB_CUDA(cudaMemsetAsync(c->values + nnz, 0, CSR_PAD * sizeof(double), stream()));
        if (T > 0) {
            unsigned nb = (unsigned)((T + 255) / 256);
            emit_entries_kernel<<<nb, 256, 0, stream()>>>(ks, flags.p, (size_t)T, m, colbits, c->colind, c->segstart,
                                                         c->rowptr, c->entry_row);
            AB_LAUNCH_CHECK();
            uint32_t Tu = (uint32_t)T;
            AB_CUDA(cudaMemcpyAsync(c->segstart + nnz, &Tu, sizeof(uint32_t), cudaMemcpyHostToDevice, stream()));
        } else {
            AB_CUDA(cudaMemsetAsync(c->segstart, 0, sizeof(uint32_t), stream()));
            AB_CUDA(cudaMemsetAsync(c->rowptr, 0, (size_t)(m + 1) * sizeof(int32_t), stream()));
        }
        // keep the permutation (needed by ab_csr_update_values)
        AB_TRY(dev_alloc((void**)&c->perm, (size_t)(T > 0 ? T : 1) * sizeof(uint32_t)));
        if (T > 0) {
            AB_CUDA(cudaMemcpyAsync(c->perm, ps, (size_t)T * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream()));
            if (nnz) {
                reduce_segments_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, stream()>>>(c->segstart, c->perm, vv.d, nnz, c->values);
                AB_LAUNCH_CHECK();
            }
        }
        AB_TRY(csr_finalize(c));
        return AB_OK;
    };
    rc = body();
    if (rc != AB_OK) { delete c; cudaEventDestroy(e0); cudaEventDestroy(e1); return rc; }
    AB_CUDA(cudaEventRecord(e1, stream()));
    AB_CUDA(cudaEventSynchronize(e1));
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    timer_add(AB_TIMER_ASSEMBLY, ms * 1e-3);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    *out_h = csr_put(c);
    return AB_OK;
}

// ---------------------------------------------------------------- entry_row / transpose
__global__ void entry_row_kernel(const int32_t* __restrict__ rowptr, int64_t m, size_t nnz, int32_t* __restrict__ er) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    // largest r with rowptr[r] <= e
    int64_t lo = 0, hi = m;   // invariant rowptr[lo] <= e < rowptr[hi]
    while (hi - lo > 1) {
        int64_t mid = (lo + hi) >> 1;
        if ((size_t)rowptr[mid] <= e) lo = mid; else hi = mid;
    }
    er[e] = (int32_t)lo;
}

// triplet -> stored-entry map, built on first use from the cached permutation (handles created
// from CSR have no permutation: the map is the identity and slot stays null)
int csr_ensure_slot(Csr* c) {
    if (c->slot || !c->perm || c->T == 0) return AB_OK;
    AB_TRY(dev_alloc((void**)&c->slot, (size_t)c->T * sizeof(int32_t)));
    if (c->nnz) {
        slot_kernel<<<(unsigned)((c->nnz + 255) / 256), 256, 0, stream()>>>(c->segstart, c->perm, (size_t)c->nnz, c->slot);
        AB_LAUNCH_CHECK();
    }
    return AB_OK;
}

int csr_ensure_entry_row(Csr* c) {
    if (c->entry_row || c->nnz == 0) return AB_OK;
    AB_TRY(dev_alloc((void**)&c->entry_row, (size_t)c->nnz * sizeof(int32_t)));
    entry_row_kernel<<<(unsigned)((c->nnz + 255) / 256), 256, 0, stream()>>>(c->rowptr, c->m, (size_t)c->nnz, c->entry_row);
    AB_LAUNCH_CHECK();
    return AB_OK;
}

__global__ void col_keys_kernel(const int32_t* __restrict__ colind, size_t nnz, uint64_t* __restrict__ keys,
                                uint32_t* __restrict__ colcnt) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    int32_t cidx = colind[e];
    keys[e] = (uint64_t)(uint32_t)cidx;
    atomicAdd(&colcnt[cidx], 1u);
}
__global__ void transpose_fill_kernel(const uint32_t* __restrict__ tperm, const int32_t* __restrict__ entry_row,
                                      const double* __restrict__ values, size_t nnz, int32_t* __restrict__ tcol,
                                      double* __restrict__ tval) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nnz) return;
    uint32_t e = tperm[k];
    tcol[k] = entry_row[e];
    tval[k] = values[e];
}
__global__ void gather_values_kernel(const uint32_t* __restrict__ tperm, const double* __restrict__ values, size_t nnz,
                                     double* __restrict__ tval) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nnz) tval[k] = values[tperm[k]];
}

// builds a fresh Csr holding A'; *tperm_out (optional) receives the A' entry -> A entry map
static int build_transpose(Csr* c, Csr** out, uint32_t** tperm_out) {
    AB_TRY(csr_ensure_entry_row(c));
    Csr* t = new Csr();
    t->m = c->n; t->n = c->m; t->nnz = c->nnz; t->T = c->nnz;
    size_t nnz = (size_t)c->nnz;
    auto body = [&]() -> int {
        AB_TRY(dev_alloc((void**)&t->rowptr, (size_t)(t->m + 1) * sizeof(int32_t)));
        AB_TRY(dev_alloc((void**)&t->colind, (nnz + CSR_PAD) * sizeof(int32_t)));
        AB_TRY(dev_alloc((void**)&t->values, (nnz + CSR_PAD) * sizeof(double)));
        AB_CUDA(cudaMemsetAsync(t->colind + nnz, 0, CSR_PAD * sizeof(int32_t), stream()));
        AB_CUDA(cudaMemsetAsync(t->values + nnz, 0, CSR_PAD * sizeof(double), stream()));
        DevBuf<uint64_t> k0, k1; DevBuf<uint32_t> p0, p1, colcnt;
        AB_TRY(k0.alloc(nnz)); AB_TRY(k1.alloc(nnz)); AB_TRY(p0.alloc(nnz)); AB_TRY(p1.alloc(nnz));
        AB_TRY(colcnt.alloc(t->m + 1));
        AB_CUDA(cudaMemsetAsync(colcnt.p, 0, (size_t)(t->m + 1) * sizeof(uint32_t), stream()));
        uint64_t* ks = k0.p; uint32_t* ps = p0.p;
        if (nnz) {
            unsigned nb = (unsigned)((nnz + 255) / 256);
            col_keys_kernel<<<nb, 256, 0, stream()>>>(c->colind, nnz, k0.p, colcnt.p);
            AB_LAUNCH_CHECK();
            AB_TRY(radix_sort_u64(k0.p, p0.p, k1.p, p1.p, nnz, bits_for(c->n > 0 ? c->n : 1), &ks, &ps));
            transpose_fill_kernel<<<nb, 256, 0, stream()>>>(ps, c->entry_row, c->values, nnz, t->colind, t->values);
            AB_LAUNCH_CHECK();
        }
        AB_TRY(scan_exclusive_u32(colcnt.p, (uint32_t*)t->rowptr, (size_t)(t->m + 1), nullptr));
        if (tperm_out) {
            AB_TRY(dev_alloc((void**)tperm_out, (nnz ? nnz : 1) * sizeof(uint32_t)));
            if (nnz) AB_CUDA(cudaMemcpyAsync(*tperm_out, ps, nnz * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream()));
        }
        AB_TRY(csr_finalize(t));
        return AB_OK;
    };
    int rc = body();
    if (rc != AB_OK) { delete t; return rc; }
    *out = t;
    return AB_OK;
}

// cached transpose of c (values refreshed through tperm after ab_csr_update_values)
int csr_ensure_transpose(Csr* c, Csr** out) {
    if (c->th) {
        Csr* t = csr_get(c->th);
        if (!t) return AB_EHANDLE;
        *out = t;
        return AB_OK;
    }
    Csr* t = nullptr;
    AB_TRY(build_transpose(c, &t, &c->tperm));
    c->th = csr_put(t);
    *out  = t;
    return AB_OK;
}

// ---------------------------------------------------------------- compress / coo_to_rows kernels
__global__ void pair_head_flags_kernel(const int64_t* __restrict__ idx2, size_t N, int rows_only,
                                       uint32_t* __restrict__ flags) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    bool head = (k == 0) || idx2[2 * k] != idx2[2 * k - 2] || (!rows_only && idx2[2 * k + 1] != idx2[2 * k - 1]);
    flags[k] = head ? 1u : 0u;
}
__global__ void compress_emit_kernel(const int64_t* __restrict__ idx2, const uint32_t* __restrict__ excl, size_t N,
                                     int64_t* __restrict__ out_idx2, uint32_t* __restrict__ segstart) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    bool head = (k == 0) || idx2[2 * k] != idx2[2 * k - 2] || idx2[2 * k + 1] != idx2[2 * k - 1];
    if (head) {
        uint32_t e = excl[k];
        out_idx2[2 * (size_t)e]     = idx2[2 * k];
        out_idx2[2 * (size_t)e + 1] = idx2[2 * k + 1];
        segstart[e] = (uint32_t)k;
    }
}
__global__ void compress_sum_kernel(const uint32_t* __restrict__ segstart, const double* __restrict__ v, size_t nout,
                                    size_t N, double* __restrict__ out_v) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nout) return;
    size_t a = segstart[e], b = (e + 1 < nout) ? segstart[e + 1] : N;
    double s = v[a];
    for (size_t k = a + 1; k < b; ++k) s += v[k];
    out_v[e] = s;
}
__global__ void compress_grad_kernel(const int64_t* __restrict__ idx2, const uint32_t* __restrict__ excl, size_t N,
                                     const double* __restrict__ grad_nv, double* __restrict__ grad_v) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N) return;
    bool head = (k == 0) || idx2[2 * k] != idx2[2 * k - 2] || idx2[2 * k + 1] != idx2[2 * k - 1];
    uint32_t e = head ? excl[k] : excl[k] - 1;
    grad_v[k] = grad_nv[e];
}
__global__ void rows_emit_kernel(const int64_t* __restrict__ idx2, const uint32_t* __restrict__ excl, size_t n,
                                 int64_t ilower, int32_t* __restrict__ rows, uint32_t* __restrict__ segstart,
                                 int32_t* __restrict__ cols) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    cols[k] = (int32_t)idx2[2 * k + 1];
    bool head = (k == 0) || idx2[2 * k] != idx2[2 * k - 2];
    if (head) {
        uint32_t e = excl[k];
        rows[e] = (int32_t)(idx2[2 * k] + ilower);
        segstart[e] = (uint32_t)k;
    }
}
__global__ void rows_ncols_kernel(const uint32_t* __restrict__ segstart, size_t nrow, size_t n, int32_t* __restrict__ ncols) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nrow) return;
    size_t b = (e + 1 < nrow) ? segstart[e + 1] : n;
    ncols[e] = (int32_t)(b - segstart[e]);
}

static int count_heads(const int64_t* d_idx2, size_t N, int rows_only, DevBuf<uint32_t>& excl, uint32_t* h_total) {
    DevBuf<uint32_t> total;
    AB_TRY(total.alloc(1));
    AB_TRY(excl.alloc(N));
    *h_total = 0;
    if (N == 0) return AB_OK;
    pair_head_flags_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream()>>>(d_idx2, N, rows_only, excl.p);
    AB_LAUNCH_CHECK();
    AB_TRY(scan_exclusive_u32(excl.p, excl.p, N, total.p));
    AB_CUDA(cudaMemcpyAsync(h_total, total.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream()));
    AB_CUDA(cudaStreamSynchronize(stream()));
    return AB_OK;
}

// ---------------------------------------------------------------- SparseAssembler
struct Assembler {
    int32_t nrow = 0;
    double  tol  = 0;
    // host staging (kept contributions, insertion order)
    std::vector<int32_t> hr, hc;
    std::vector<double>  hv;
    // device-resident part (earlier flushes)
    int32_t* dr = nullptr; int32_t* dc = nullptr; double* dv = nullptr;
    size_t dn = 0, dcap = 0;
    ~Assembler() { dev_free(dr); dev_free(dc); dev_free(dv); }
};
static std::map<int32_t, Assembler*> g_asm;
static std::mutex g_asm_mu;

static int asm_reserve(Assembler* a, size_t need) {
    if (need <= a->dcap) return AB_OK;
    size_t cap = a->dcap ? a->dcap : 1 << 16;
    while (cap < need) cap *= 2;
    int32_t *nr, *nc; double* nv;
    AB_TRY(dev_alloc((void**)&nr, cap * sizeof(int32_t)));
    AB_TRY(dev_alloc((void**)&nc, cap * sizeof(int32_t)));
    AB_TRY(dev_alloc((void**)&nv, cap * sizeof(double)));
    if (a->dn) {
        AB_CUDA(cudaMemcpyAsync(nr, a->dr, a->dn * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream()));
        AB_CUDA(cudaMemcpyAsync(nc, a->dc, a->dn * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream()));
        AB_CUDA(cudaMemcpyAsync(nv, a->dv, a->dn * sizeof(double), cudaMemcpyDeviceToDevice, stream()));
    }
    dev_free(a->dr); dev_free(a->dc); dev_free(a->dv);
    a->dr = nr; a->dc = nc; a->dv = nv; a->dcap = cap;
    return AB_OK;
}
static int asm_flush(Assembler* a) {
    size_t k = a->hv.size();
    if (!k) return AB_OK;
    AB_TRY(asm_reserve(a, a->dn + k));
    AB_CUDA(cudaMemcpyAsync(a->dr + a->dn, a->hr.data(), k * sizeof(int32_t), cudaMemcpyHostToDevice, stream()));
    AB_CUDA(cudaMemcpyAsync(a->dc + a->dn, a->hc.data(), k * sizeof(int32_t), cudaMemcpyHostToDevice, stream()));
    AB_CUDA(cudaMemcpyAsync(a->dv + a->dn, a->hv.data(), k * sizeof(double), cudaMemcpyHostToDevice, stream()));
    AB_CUDA(cudaStreamSynchronize(stream()));
    a->dn += k;
    a->hr.clear(); a->hc.clear(); a->hv.clear();
    return AB_OK;
}

__global__ void asm_filter_flags_kernel(const int32_t* __restrict__ rows, const double* __restrict__ vals, size_t k,
                                        double tol, int32_t nrow, uint32_t* __restrict__ flags, int* __restrict__ err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    bool keep = fabs(vals[i]) > tol;
    if (keep && (rows[i] < 1 || rows[i] > nrow)) { atomicExch(err, 1); keep = false; }
    flags[i] = keep ? 1u : 0u;
}
__global__ void asm_compact_kernel(const int32_t* __restrict__ rows, const int32_t* __restrict__ cols,
                                   const double* __restrict__ vals, size_t k, double tol, const uint32_t* __restrict__ excl,
                                   int32_t* __restrict__ orow, int32_t* __restrict__ ocol, double* __restrict__ oval) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    if (fabs(vals[i]) > tol) {
        uint32_t p = excl[i];
        orow[p] = rows[i]; ocol[p] = cols[i]; oval[p] = vals[i];
    }
}
__global__ void asm_row_keys_kernel(const int32_t* __restrict__ rows, size_t k, uint64_t* __restrict__ keys) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) keys[i] = (uint64_t)(uint32_t)(rows[i] - 1);
}
__global__ void asm_gather_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ perm,
                                  const int32_t* __restrict__ cols, const double* __restrict__ vals, size_t k,
                                  int32_t* __restrict__ orow, int32_t* __restrict__ ocol, double* __restrict__ oval) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    uint32_t p = perm[i];
    orow[i] = (int32_t)keys[i] + 1;
    ocol[i] = cols[p];
    oval[i] = vals[p];
}

}  // namespace ab

using namespace ab;

extern "C" {

int ab_csr_from_coo(int64_t T, const int64_t* ii, const int64_t* jj, const double* vv, int64_t m, int64_t n,
                    int base, int64_t* out_h) {
    return csr_from_coo_impl<int64_t>(T, ii, jj, vv, m, n, base, out_h);
}
int ab_csr_from_coo32(int64_t T, const int32_t* ii, const int32_t* jj, const double* vv, int64_t m, int64_t n,
                      int base, int64_t* out_h) {
    return csr_from_coo_impl<int32_t>(T, ii, jj, vv, m, n, base, out_h);
}

int ab_csr_create(int64_t m, int64_t n, int64_t nnz, const int32_t* rowptr_, const int32_t* colind_,
                  const double* values_, int64_t* out_h) {
    AB_TRY(ensure_device());
    if (!out_h || !rowptr_ || (nnz > 0 && (!colind_ || !values_))) return fail(AB_EINVAL, "null argument");
    if (m < 0 || n < 0 || nnz < 0 || nnz >= INT32_MAX || m >= INT32_MAX || n >= INT32_MAX)
        return fail(AB_EUNSUPPORTED, "size outside int32 local index space");
    std::lock_guard<std::mutex> lk(big_lock());
    Csr* c = new Csr();
    c->m = m; c->n = n; c->nnz = nnz; c->T = nnz;
    auto body = [&]() -> int {
        AB_TRY(dev_alloc((void**)&c->rowptr, (size_t)(m + 1) * sizeof(int32_t)));
        AB_TRY(dev_alloc((void**)&c->colind, ((size_t)nnz + CSR_PAD) * sizeof(int32_t)));
        AB_TRY(dev_alloc((void**)&c->values, ((size_t)nnz + CSR_PAD) * sizeof(double)));
        AB_CUDA(cudaMemcpyAsync(c->rowptr, rowptr_, (size_t)(m + 1) * sizeof(int32_t), cudaMemcpyDefault, stream()));
        if (nnz) {
            AB_CUDA(cudaMemcpyAsync(c->colind, colind_, (size_t)nnz * sizeof(int32_t), cudaMemcpyDefault, stream()));
            AB_CUDA(cudaMemcpyAsync(c->values, values_, (size_t)nnz * sizeof(double), cudaMemcpyDefault, stream()));
        }
        AB_CUDA(cudaMemsetAsync(c->colind + nnz, 0, CSR_PAD * sizeof(int32_t), stream()));
        AB_CUDA(cudaMemsetAsync(c->values + nnz, 0, CSR_PAD * sizeof(double), stream()));
        AB_TRY(csr_finalize(c));
        return AB_OK;
    };
    int rc = body();
    if (rc != AB_OK) { delete c; return rc; }
    *out_h = csr_put(c);
    return AB_OK;
}

int ab_csr_export(int64_t h, int32_t* rowptr, int32_t* colind, double* values) {
    AB_TRY(ensure_device());
    Csr* c = csr_get(h);
    if (!c) return AB_EHANDLE;
    std::lock_guard<std::mutex> lk(big_lock());
    if (rowptr) AB_CUDA(cudaMemcpyAsync(rowptr, c->rowptr, (size_t)(c->m + 1) * sizeof(int32_t), cudaMemcpyDefault, stream()));
    if (colind && c->nnz) AB_CUDA(cudaMemcpyAsync(colind, c->colind, (size_t)c->nnz * sizeof(int32_t), cudaMemcpyDefault, stream()));
    if (values && c->nnz) AB_CUDA(cudaMemcpyAsync(values, c->values, (size_t)c->nnz * sizeof(double), cudaMemcpyDefault, stream()));
    AB_CUDA(cudaStreamSynchronize(stream()));
    return AB_OK;
}

__global__ void iota_i32_kernel(int32_t* p, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int32_t)i;
}

int ab_csr_export_map(int64_t h, int32_t* slot) {
    AB_TRY(ensure_device());
    Csr* c = csr_get(h);
    if (!c) return AB_EHANDLE;
    if (!slot) return fail(AB_EINVAL, "slot is null");
    std::lock_guard<std::mutex> lk(big_lock());
    if (c->T == 0) return AB_OK;
    AB_TRY(csr_ensure_slot(c));
    if (c->slot) {
        AB_CUDA(cudaMemcpyAsync(slot, c->slot, (size_t)c->T * sizeof(int32_t), cudaMemcpyDefault, stream()));
    } else {
        Out<int32_t> o;
        AB_TRY(o.bind(slot, c->T));
        iota_i32_kernel<<<(unsigned)((c->T + 255) / 256), 256, 0, stream()>>>(o.d, (size_t)c->T);
        AB_LAUNCH_CHECK();
        AB_TRY(o.commit());
    }
    AB_CUDA(cudaStreamSynchronize(stream()));
    return AB_OK;
}

int ab_csr_update_values(int64_t h, const double* vv_) {
    AB_TRY(ensure_device());
    Csr* c = csr_get(h);
    if (!c) return AB_EHANDLE;
    if (c->T > 0 && !vv_) return fail(AB_EINVAL, "vv is null");
    std::lock_guard<std::mutex> lk(big_lock());
    In<double> vv;
    AB_TRY(vv.bind(vv_, c->T));
    if (c->nnz) {
        if (c->perm) {
            reduce_segments_kernel<<<(unsigned)((c->nnz + 255) / 256), 256, 0, stream()>>>(c->segstart, c->perm, vv.d, (size_t)c->nnz, c->values);
            AB_LAUNCH_CHECK();
        } else {
            AB_CUDA(cudaMemcpyAsync(c->values, vv.d, (size_t)c->nnz * sizeof(double), cudaMemcpyDeviceToDevice, stream()));
        }
    }
    c->dinv_valid = false;
    ++c->values_version;
    if (c->th) {   // refresh cached transpose values
        Csr* t = csr_get(c->th);
        if (t && c->nnz) {
            gather_values_kernel<<<(unsigned)((c->nnz + 255) / 256), 256, 0, stream()>>>(c->tperm, c->values, (size_t)c->nnz, t->values);
            AB_LAUNCH_CHECK();
            t->dinv_valid = false;
            ++t->values_version;
        }
    }
    if (vv.tmp.p) AB_CUDA(cudaStreamSynchronize(stream()));
    return AB_OK;
}

int ab_csr_transpose(int64_t h, int64_t* out_h) {
    AB_TRY(ensure_device());
    Csr* c = csr_get(h);
    if (!c) return AB_EHANDLE;
    if (!out_h) return fail(AB_EINVAL, "out_h is null");
    std::lock_guard<std::mutex> lk(big_lock());
    Csr* t = nullptr;
    AB_TRY(build_transpose(c, &t, nullptr));
    *out_h = csr_put(t);
    return AB_OK;
}

// ------------------------------------------------------------------ tf.sparse.reorder (SparseTensor ctor / find)
__global__ void reorder_gather_kernel(const uint32_t* __restrict__ perm, size_t T, const int64_t* __restrict__ ii,
                                      const int64_t* __restrict__ jj, const double* __restrict__ vv,
                                      int64_t* __restrict__ oi, int64_t* __restrict__ oj, double* __restrict__ ov,
                                      int64_t* __restrict__ operm) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= T) return;
    const uint32_t t = perm[k];
    if (oi) oi[k] = ii[t];
    if (oj) oj[k] = jj[t];
    if (ov) ov[k] = vv[t];
    if (operm) operm[k] = (int64_t)t;
}

// Stable row-major sort of the triplets, duplicates KEPT (src/sparse.jl:62-76: tf.sparse.reorder in the
// constructor; find() then hands (ii, jj, vv) in this order to every op).  Any output may be NULL;
// perm[k] = input position of output k (0-based).
int ab_coo_reorder(int64_t T, const int64_t* ii_, const int64_t* jj_, const double* vv_, int64_t m, int64_t n, int base,
                   int64_t* oii_, int64_t* ojj_, double* ovv_, int64_t* perm_) {
    AB_TRY(ensure_device());
    if (T < 0 || m < 0 || n < 0) return fail(AB_EINVAL, "negative size");
    if (T == 0) return AB_OK;
    if (!ii_ || !jj_ || (ovv_ && !vv_)) return fail(AB_EINVAL, "null triplet array");
    if (m >= INT32_MAX || n >= INT32_MAX) return fail(AB_EUNSUPPORTED, "matrix dimension exceeds int32 local index space");
    if (T >= (int64_t)UINT32_MAX) return fail(AB_EUNSUPPORTED, "more than 2^32-1 triplets in one call");
    std::lock_guard<std::mutex> lk(big_lock());
    In<int64_t> ii, jj; In<double> vv;
    AB_TRY(ii.bind(ii_, T)); AB_TRY(jj.bind(jj_, T)); AB_TRY(vv.bind(vv_, T));
    Out<int64_t> oi, oj, op; Out<double> ov;
    AB_TRY(oi.bind(oii_, T)); AB_TRY(oj.bind(ojj_, T)); AB_TRY(ov.bind(ovv_, T)); AB_TRY(op.bind(perm_, T));
    const int colbits = bits_for(n > 0 ? n : 1), rowbits = bits_for(m > 0 ? m : 1);
    DevBuf<uint64_t> k0, k1; DevBuf<uint32_t> p0, p1; DevBuf<int> err;
    AB_TRY(k0.alloc(T)); AB_TRY(k1.alloc(T)); AB_TRY(p0.alloc(T)); AB_TRY(p1.alloc(T)); AB_TRY(err.alloc(1));
    AB_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), stream()));
    const unsigned nb = (unsigned)((T + 255) / 256);
    make_keys_kernel<int64_t><<<nb, 256, 0, stream()>>>(ii.d, jj.d, (size_t)T, m, n, base, colbits, k0.p, err.p);
    AB_LAUNCH_CHECK();
    uint64_t* ks; uint32_t* ps;
    AB_TRY(radix_sort_u64(k0.p, p0.p, k1.p, p1.p, (size_t)T, colbits + rowbits, &ks, &ps));
    int herr = 0;
    AB_CUDA(cudaMemcpyAsync(&herr, err.p, sizeof(int), cudaMemcpyDeviceToHost, stream()));
    AB_CUDA(cudaStreamSynchronize(stream()));
    if (herr) return fail(AB_ERANGE, "triplet index outside [%d,%lld]x[%d,%lld]", base, (long long)(m + base - 1), base,
                          (long long)(n + base - 1));
    reorder_gather_kernel<<<nb, 256, 0, stream()>>>(ps, (size_t)T, ii.d, jj.d, vv.d, oi.d, oj.d, ov.d, op.d);
    AB_LAUNCH_CHECK();
    AB_TRY(oi.commit()); AB_TRY(oj.commit()); AB_TRY(ov.commit()); AB_TRY(op.commit());
    AB_CUDA(cudaStreamSynchronize(stream()));
    return AB_OK;
}

// ------------------------------------------------------------------ compress
int ab_compress_count(int64_t N, const int64_t* idx2_, int64_t* nout) {
    AB_TRY(ensure_device());
    if (N < 0 || !nout || (N > 0 && !idx2_)) return fail(AB_EINVAL, "bad argument");
    if (N >= (int64_t)UINT32_MAX) return fail(AB_EUNSUPPORTED, "N exceeds 32-bit positions");
    std::lock_guard<std::mutex> lk(big_lock());
    In<int64_t> idx2;
    AB_TRY(idx2.bind(idx2_, 2 * (size_t)N));
    DevBuf<uint32_t> excl; uint32_t tot = 0;
    AB_TRY(count_heads(idx2.d, (size_t)N, 0, excl, &tot));
    *nout = tot;
    return AB_OK;
}

int ab_compress(int64_t N, const int64_t* idx2_, const double* v_, int64_t* out_idx2_, double* out_v_) {
    AB_TRY(ensure_device());
    if (N < 0 || (N > 0 && (!idx2_ || !v_ || !out_idx2_ || !out_v_))) return fail(AB_EINVAL, "bad argument");
    if (N >= (int64_t)UINT32_MAX) return fail(AB_EUNSUPPORTED, "N exceeds 32-bit positions");
    if (N == 0) return AB_OK;
    std::lock_guard<std::mutex> lk(big_lock());
    In<int64_t> idx2; In<double> v;
    AB_TRY(idx2.bind(idx2_, 2 * (size_t)N)); AB_TRY(v.bind(v_, N));
    DevBuf<uint32_t> excl, seg; uint32_t nout = 0;
    AB_TRY(count_heads(idx2.d, (size_t)N, 0, excl, &nout));
    Out<int64_t> oi; Out<double> ov;
    AB_TRY(oi.bind(out_idx2_, 2 * (size_t)nout)); AB_TRY(ov.bind(out_v_, nout));
    AB_TRY(seg.alloc(nout));
    unsigned nb = (unsigned)((N + 255) / 256);
    compress_emit_kernel<<<nb, 256, 0, stream()>>>(idx2.d, excl.p, (size_t)N, oi.d, seg.p);
    AB_LAUNCH_CHECK();
    compress_sum_kernel<<<(nout + 255) / 256, 256, 0, stream()>>>(seg.p, v.d, nout, (size_t)N, ov.d);
    AB_LAUNCH_CHECK();
    AB_TRY(oi.commit()); AB_TRY(ov.commit());
    AB_CUDA(cudaStreamSynchronize(stream()));
    return AB_OK;
}

int ab_compress_grad(int64_t N, const int64_t* idx2_, const double* grad_nv_, double* grad_v_) {
    AB_TRY(ensure_device());
    if (N < 0 || (N > 0 && (!idx2_ || !grad_nv_ || !grad_v_))) return fail(AB_EINVAL, "bad argument");
    if (N == 0) return AB_OK;
    std::lock_guard<std::mutex> lk(big_lock());
    In<int64_t> idx2;
    AB_TRY(idx2.bind(idx2_, 2 * (size_t)N));
    DevBuf<uint32_t> excl; uint32_t nout = 0;
    AB_TRY(count_heads(idx2.d, (size_t)N, 0, excl, &nout));
    In<double> g; Out<double> o;
    AB_TRY(g.bind(grad_nv_, nout)); AB_TRY(o.bind(grad_v_, N));
    compress_grad_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream()>>>(idx2.d, excl.p, (size_t)N, g.d, o.d);
    AB_LAUNCH_CHECK();
    AB_TRY(o.commit());
    AB_CUDA(cudaStreamSynchronize(stream()));
    return AB_OK;
}

// ------------------------------------------------------------------ mpi_create_matrix
int ab_coo_to_rows_count(int64_t n, const int64_t* idx2_, int64_t* nrow) {
    AB_TRY(ensure_device());
    if (n < 0 || !nrow || (n > 0 && !idx2_)) return fail(AB_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(big_lock());
    In<int64_t> idx2;
    AB_TRY(idx2.bind(idx2_, 2 * (size_t)n));
    DevBuf<uint32_t> excl; uint32_t tot = 0;
    AB_TRY(count_heads(idx2.d, (size_t)n, 1, excl, &tot));
    *nrow = tot;
    return AB_OK;
}
int ab_coo_to_rows(int64_t n, const int64_t* idx2_, int64_t ilower, int32_t* rows_, int32_t* ncols_, int32_t* cols_) {
    AB_TRY(ensure_device());
    if (n < 0 || (n > 0 && (!idx2_ || !rows_ || !ncols_ || !cols_))) return fail(AB_EINVAL, "bad argument");
    if (n == 0) return AB_OK;
    std::lock_guard<std::mutex> lk(big_lock());
    In<int64_t> idx2;
    AB_TRY(idx2.bind(idx2_, 2 * (size_t)n));
    DevBuf<uint32_t> excl, seg; uint32_t nrow = 0;
    AB_TRY(count_heads(idx2.d, (size_t)n, 1, excl, &nrow));
    Out<int32_t> r, nc, cc;
    AB_TRY(r.bind(rows_, nrow)); AB_TRY(nc.bind(ncols_, nrow)); AB_TRY(cc.bind(cols_, n));
    AB_TRY(seg.alloc(nrow));
    rows_emit_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream()>>>(idx2.d, excl.p, (size_t)n, ilower, r.d, seg.p, cc.d);
    AB_LAUNCH_CHECK();
    rows_ncols_kernel<<<(nrow + 255) / 256, 256, 0, stream()>>>(seg.p, nrow, (size_t)n, nc.d);
    AB_LAUNCH_CHECK();
    AB_TRY(r.commit()); AB_TRY(nc.commit()); AB_TRY(cc.commit());
    AB_CUDA(cudaStreamSynchronize(stream()));
    return AB_OK;
}

// ------------------------------------------------------------------ SparseAssembler
int ab_asm_create(int32_t h, int32_t nrow, double tol) {
    if (nrow < 0) return fail(AB_EINVAL, "nrow < 0");
    std::lock_guard<std::mutex> lk(g_asm_mu);
    if (g_asm.count(h)) return AB_OK;   // existing handle is returned as is (Impl.cpp:61-78)
    Assembler* a = new Assembler();
    a->nrow = nrow; a->tol = tol;
    g_asm[h] = a;
    return AB_OK;
}

static Assembler* asm_get(int32_t h) {
    auto it = g_asm.find(h);
    if (it == g_asm.end()) { set_error("unknown assembler handle %d", h); return nullptr; }
    return it->second;
}

int ab_asm_add(int32_t h, int32_t row, const int32_t* cols, const double* vals, int32_t k) {
    if (k < 0 || (k > 0 && (!cols || !vals))) return fail(AB_EINVAL, "bad argument");
    std::lock_guard<std::mutex> lk(g_asm_mu);
    Assembler* a = asm_get(h);
    if (!a) return AB_EHANDLE;
    if (k > 0 && (is_device_ptr(cols) || is_device_ptr(vals)))
        return fail(AB_EINVAL, "ab_asm_add takes host arrays; use ab_asm_add_coo for device data");
    for (int32_t i = 0; i < k; ++i) {
        if (fabs(vals[i]) > a->tol) {   // strict, per contribution (Impl.cpp:16)
            if (row < 1 || row > a->nrow)
                return fail(AB_ERANGE, "row_id out of the bounds %d > %d", row, a->nrow);
            a->hr.push_back(row); a->hc.push_back(cols[i]); a->hv.push_back(vals[i]);
        }
    }
    return AB_OK;
}

int ab_asm_add_coo(int32_t h, const int32_t* rows_, const int32_t* cols_, const double* vals_, int64_t k) {
    if (k < 0 || (k > 0 && (!rows_ || !cols_ || !vals_))) return fail(AB_EINVAL, "bad argument");
    if (k == 0) return AB_OK;
    if (k >= (int64_t)UINT32_MAX) return fail(AB_EUNSUPPORTED, "k exceeds 32-bit positions");
    AB_TRY(ensure_device());
    std::lock_guard<std::mutex> lk(g_asm_mu);
    std::lock_guard<std::mutex> lk2(big_lock());
    Assembler* a = asm_get(h);
    if (!a) return AB_EHANDLE;
    AB_TRY(asm_flush(a));   // keep insertion order
    In<int32_t> r, c; In<double> v;
    AB_TRY(r.bind(rows_, k)); AB_TRY(c.bind(cols_, k)); AB_TRY(v.bind(vals_, k));
    DevBuf<uint32_t> flags, total; DevBuf<int> err;
    AB_TRY(flags.alloc(k)); AB_TRY(total.alloc(1)); AB_TRY(err.alloc(1));
    AB_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), stream()));
    unsigned nb = (unsigned)((k + 255) / 256);
    asm_filter_flags_kernel<<<nb, 256, 0, stream()>>>(r.d, v.d, (size_t)k, a->tol, a->nrow, flags.p, err.p);
    AB_LAUNCH_CHECK();
    AB_TRY(scan_exclusive_u32(flags.p, flags.p, (size_t)k, total.p));
    uint32_t kept = 0; int herr = 0;
    AB_CUDA(cudaMemcpyAsync(&kept, total.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream()));
    AB_CUDA(cudaMemcpyAsync(&herr, err.p, sizeof(int), cudaMemcpyDeviceToHost, stream()));
    AB_CUDA(cudaStreamSynchronize(stream()));
    if (herr) return fail(AB_ERANGE, "row_id out of the bounds [1,%d]", a->nrow);
    AB_TRY(asm_reserve(a, a->dn + kept));
    asm_compact_kernel<<<nb, 256, 0, stream()>>>(r.d, c.d, v.d, (size_t)k, a->tol, flags.p, a->dr + a->dn, a->dc + a->dn, a->dv + a->dn);
    AB_LAUNCH_CHECK();
    a->dn += kept;
    AB_CUDA(cudaStreamSynchronize(stream()));
    return AB_OK;
}

int64_t ab_asm_nnz(int32_t h) {
    std::lock_guard<std::mutex> lk(g_asm_mu);
    Assembler* a = asm_get(h);
    if (!a) return AB_EHANDLE;
    return (int64_t)(a->dn + a->hv.size());
}

int ab_asm_destroy(int32_t h) {
    std::lock_guard<std::mutex> lk(g_asm_mu);
    auto it = g_asm.find(h);
    if (it == g_asm.end()) return fail(AB_EHANDLE, "unknown assembler handle %d", h);
    delete it->second;
    g_asm.erase(it);
    return AB_OK;
}

int ab_asm_copy(int32_t h, int32_t* rows_, int32_t* cols_, double* vals_) {
    AB_TRY(ensure_device());
    std::lock_guard<std::mutex> lk(g_asm_mu);
    std::lock_guard<std::mutex> lk2(big_lock());
    Assembler* a = asm_get(h);
    if (!a) return AB_EHANDLE;
    AB_TRY(asm_flush(a));
    size_t k = a->dn;
    if (k) {
        if (!rows_ || !cols_ || !vals_) return fail(AB_EINVAL, "null output");
        Out<int32_t> r, c; Out<double> v;
        AB_TRY(r.bind(rows_, k)); AB_TRY(c.bind(cols_, k)); AB_TRY(v.bind(vals_, k));
        DevBuf<uint64_t> k0, k1; DevBuf<uint32_t> p0, p1;
        AB_TRY(k0.alloc(k)); AB_TRY(k1.alloc(k)); AB_TRY(p0.alloc(k)); AB_TRY(p1.alloc(k));
        unsigned nb = (unsigned)((k + 255) / 256);
        asm_row_keys_kernel<<<nb, 256, 0, stream()>>>(a->dr, k, k0.p);
        AB_LAUNCH_CHECK();
        uint64_t* ks; uint32_t* ps;
        AB_TRY(radix_sort_u64(k0.p, p0.p, k1.p, p1.p, k, bits_for(a->nrow > 0 ? a->nrow : 1), &ks, &ps));
        asm_gather_kernel<<<nb, 256, 0, stream()>>>(ks, ps, a->dc, a->dv, k, r.d, c.d, v.d);
        AB_LAUNCH_CHECK();
        AB_TRY(r.commit()); AB_TRY(c.commit()); AB_TRY(v.commit());
        AB_CUDA(cudaStreamSynchronize(stream()));
    }
    delete a;           // the handle is destroyed by the dump (Impl.cpp:107-118)
    g_asm.erase(h);
    return AB_OK;
}

}  // extern "C"
