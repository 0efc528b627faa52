// structure.cu — structural edits that sit between assembly and solve in every ADCME FEM script
//  and the remaining SparseTensor algebra (N4), all on the device:
//
//   zero_out_row            deps/CustomOps/ZeroOutRow/ZeroOutRow.h:28-57        (src/sparse.jl:790-797)
//   sparse_scatter_update   deps/CustomOps/SparseScatterUpdate/SparseScatterUpdate.h:20-60 (sparse.jl:340-364)
//   sparse_indexing         deps/CustomOps/SparseIndexing/SparseIndexing.h:23-83 (sparse.jl:303-327)
//   sparse_sparse_mat_mul   deps/CustomOps/SparseMatMul/SparseMatMul.h:11-52    (sparse.jl:579-595)
//   diag_sparse / sparse_diag fast paths: the same header, :34-52 (ab_coo_scale + assembly)
//   A + B                   src/sparse.jl:203-218 -> tf.sparse.add
//
// The reference ops return triplet lists whose length is only known after the op ran (they size
// allocate_output from op-private state: IJV.get_size(), ZeroOutRow::forward's return value).
// Here such a result is parked on the device under a "triplet handle": ab_trip_size -> caller
// allocates -> ab_trip_fetch (frees the handle), or it is assembled straight into a CSR handle
// with ab_csr_from_trip without leaving the GPU.
//
// Kernels are streaming flag -> scan -> compact passes (HBM-bound integer/byte work, one read of the
// inputs and one write of the outputs each); SpGEMM is expand -> sort -> segmented reduce on top of
// the assembly pipeline of assemble.cu, which fixes the summation order of every C(i,j) to
// k-ascending — the order Eigen's column-major product uses (SparseMatMul.h:31).
#include "common.cuh"
#include <map>

namespace ab {

// ---------------------------------------------------------------- triplet handles
struct Trip {
    int64_t  n  = 0;
    int64_t* ii = nullptr;
    int64_t* jj = nullptr;
    double*  vv = nullptr;
    ~Trip() { dev_free(ii); dev_free(jj); dev_free(vv); }
};
static std::map<int64_t, Trip*> g_trip;
static std::mutex g_trip_mu;
static int64_t g_next_trip = 1;

static int64_t trip_put(Trip* t) {
    std::lock_guard<std::mutex> lk(g_trip_mu);
    int64_t h = g_next_trip++;
    g_trip[h] = t;
    return h;
}
static Trip* trip_get(int64_t h) {
    std::lock_guard<std::mutex> lk(g_trip_mu);
    auto it = g_trip.find(h);
    if (it == g_trip.end()) { set_error("unknown triplet handle %lld", (long long)h); return nullptr; }
    return it->second;
}
static Trip* trip_take(int64_t h) {
    std::lock_guard<std::mutex> lk(g_trip_mu);
    auto it = g_trip.find(h);
    if (it == g_trip.end()) { set_error("unknown triplet handle %lld", (long long)h); return nullptr; }
    Trip* t = it->second;
    g_trip.erase(it);
    return t;
}
static int trip_alloc(Trip* t, int64_t n) {
    t->n = n;
    size_t c = (size_t)(n > 0 ? n : 1);
    AB_TRY(dev_alloc((void**)&t->ii, c * sizeof(int64_t)));
    AB_TRY(dev_alloc((void**)&t->jj, c * sizeof(int64_t)));
    AB_TRY(dev_alloc((void**)&t->vv, c * sizeof(double)));
    return AB_OK;
}

static inline unsigned nblk(size_t n) { return (unsigned)((n + 255) / 256); }
static int bits_for(int64_t dim) {
    int b = 1;
    while (b < 63 && ((int64_t)1 << b) < dim) ++b;
    return b;
}

static int fetch_u32(const uint32_t* d, uint32_t* h) {
    AB_CUDA(cudaMemcpyAsync(h, d, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream()));
    AB_CUDA(cudaStreamSynchronize(stream()));
    return AB_OK;
}
static int fetch_err(const int* d, int* h) {
    AB_CUDA(cudaMemcpyAsync(h, d, sizeof(int), cudaMemcpyDeviceToHost, stream()));
    AB_CUDA(cudaStreamSynchronize(stream()));
    return AB_OK;
}

// ---------------------------------------------------------------- shared small kernels
// mask[v - base] = 1 for every listed index; out-of-range values raise *err
__global__ void mark_mask_kernel(const int64_t* __restrict__ idx, size_t k, int64_t base, int64_t dim,
                                 uint8_t* __restrict__ mask, int* __restrict__ err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    const int64_t v = idx[i] - base;
    if (v < 0 || v >= dim) { atomicExch(err, 1); return; }
    mask[v] = 1;
}
// map[v - 1] = LAST position p with idx[p] == v (std::map overwrite semantics, SparseIndexing.h:27-32)
__global__ void mark_map_kernel(const int64_t* __restrict__ idx, size_t k, int64_t dim, int32_t* __restrict__ map,
                                int* __restrict__ err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    const int64_t v = idx[i] - 1;
    if (v < 0 || v >= dim) { atomicExch(err, 1); return; }
    atomicMax(&map[v], (int32_t)i);
}
__global__ void max_i64_kernel(const int64_t* __restrict__ v, size_t k, unsigned long long* __restrict__ out,
                               int* __restrict__ err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    if (v[i] < 1) { atomicExch(err, 1); return; }
    atomicMax(out, (unsigned long long)v[i]);
}

// ================================================================ zero_out_row
// keep[i] = 1 unless the (0-based) row of entry i is in {bd - 1}   (ZeroOutRow.h:28-38)
__global__ void zor_flags_kernel(const int64_t* __restrict__ idx2, size_t N, const uint8_t* __restrict__ mask,
                                 int64_t masklen, uint32_t* __restrict__ keep) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int64_t r = idx2[2 * i];
    keep[i] = (r >= 0 && r < masklen && mask[r]) ? 0u : 1u;
}
__global__ void zor_compact_kernel(const int64_t* __restrict__ idx2, const double* __restrict__ v, size_t N,
                                   const uint8_t* __restrict__ mask, int64_t masklen,
                                   const uint32_t* __restrict__ excl, int64_t* __restrict__ oi,
                                   int64_t* __restrict__ oj, double* __restrict__ ov) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int64_t r = idx2[2 * i];
    if (r >= 0 && r < masklen && mask[r]) return;
    const uint32_t p = excl[i];
    oi[p] = r; oj[p] = idx2[2 * i + 1]; ov[p] = v[i];
}
__global__ void zor_grad_kernel(const int64_t* __restrict__ idx2, size_t N, const uint8_t* __restrict__ mask,
                                int64_t masklen, const uint32_t* __restrict__ excl,
                                const double* __restrict__ grad_ov, double* __restrict__ grad_v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int64_t r = idx2[2 * i];
    const bool drop = (r >= 0 && r < masklen && mask[r]);
    grad_v[i] = drop ? 0.0 : grad_ov[excl[i]];   // the reference leaves dropped slots unwritten; 0 here
}

// row mask of {bd - 1}, sized by max(bd); masklen = 0 when bd is empty
static int zor_mask(const int64_t* bd_d, int64_t nbd, DevBuf<uint8_t>& mask, int64_t* masklen) {
    *masklen = 0;
    if (nbd == 0) return mask.alloc(1);
    DevBuf<unsigned long long> mx; DevBuf<int> err;
    AB_TRY(mx.alloc(1)); AB_TRY(err.alloc(1));
    AB_CUDA(cudaMemsetAsync(mx.p, 0, sizeof(unsigned long long), stream()));
    AB_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), stream()));
    max_i64_kernel<<<nblk(nbd), 256, 0, stream()>>>(bd_d, (size_t)nbd, mx.p, err.p);
    AB_LAUNCH_CHECK();
    unsigned long long hmx = 0; int herr = 0;
    AB_CUDA(cudaMemcpyAsync(&hmx, mx.p, sizeof(hmx), cudaMemcpyDeviceToHost, stream()));
    AB_TRY(fetch_err(err.p, &herr));
    if (herr) return fail(AB_ERANGE, "zero_out_row: bd holds an index < 1");
    if (hmx >= (unsigned long long)INT32_MAX) return fail(AB_EUNSUPPORTED, "zero_out_row: bd index exceeds int32");
    *masklen = (int64_t)hmx;
    AB_TRY(mask.alloc((size_t)hmx));
    AB_CUDA(cudaMemsetAsync(mask.p, 0, (size_t)hmx, stream()));
    mark_mask_kernel<<<nblk(nbd), 256, 0, stream()>>>(bd_d, (size_t)nbd, 1, (int64_t)hmx, mask.p, err.p);
    AB_LAUNCH_CHECK();
    return AB_OK;
}

// ================================================================ sparse_scatter_update
// keep[i] = !(rmask[oii-1] && cmask[ojj-1])      (SparseScatterUpdate.h:27-33)
__global__ void ssu_flags_kernel(const int64_t* __restrict__ oii, const int64_t* __restrict__ ojj, size_t on,
                                 const uint8_t* __restrict__ rmask, const uint8_t* __restrict__ cmask, int64_t m,
                                 int64_t n, uint32_t* __restrict__ keep, int* __restrict__ err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= on) return;
    const int64_t r = oii[i] - 1, c = ojj[i] - 1;
    if (r < 0 || r >= m || c < 0 || c >= n) { atomicExch(err, 1); keep[i] = 0; return; }
    keep[i] = (rmask[r] && cmask[c]) ? 0u : 1u;
}
__global__ void ssu_compact_kernel(const int64_t* __restrict__ oii, const int64_t* __restrict__ ojj,
                                   const double* __restrict__ ovv, size_t on, const uint32_t* __restrict__ keep_excl,
                                   const uint8_t* __restrict__ rmask, const uint8_t* __restrict__ cmask,
                                   int64_t* __restrict__ oi, int64_t* __restrict__ oj, double* __restrict__ ov) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= on) return;
    const int64_t r = oii[i] - 1, c = ojj[i] - 1;
    if (rmask[r] && cmask[c]) return;
    const uint32_t p = keep_excl[i];
    oi[p] = oii[i]; oj[p] = ojj[i]; ov[p] = ovv[i];
}
// appended block entries: (ii[uii-1], jj[ujj-1], uvv)   (SparseScatterUpdate.h:34-36)
__global__ void ssu_append_kernel(const int64_t* __restrict__ uii, const int64_t* __restrict__ ujj,
                                  const double* __restrict__ uvv, size_t un, const int64_t* __restrict__ ii,
                                  int64_t ni, const int64_t* __restrict__ jj, int64_t nj, int64_t* __restrict__ oi,
                                  int64_t* __restrict__ oj, double* __restrict__ ov, int* __restrict__ err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= un) return;
    const int64_t a = uii[i] - 1, b = ujj[i] - 1;
    if (a < 0 || a >= ni || b < 0 || b >= nj) { atomicExch(err, 1); oi[i] = 1; oj[i] = 1; ov[i] = 0.0; return; }
    oi[i] = ii[a]; oj[i] = jj[b]; ov[i] = uvv[i];
}
__global__ void ssu_grad_kernel(const int64_t* __restrict__ oii, const int64_t* __restrict__ ojj, size_t on,
                                const uint32_t* __restrict__ keep_excl, const uint8_t* __restrict__ rmask,
                                const uint8_t* __restrict__ cmask, const double* __restrict__ grad_out,
                                double* __restrict__ grad_ovv) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= on) return;
    const int64_t r = oii[i] - 1, c = ojj[i] - 1;
    grad_ovv[i] = (rmask[r] && cmask[c]) ? 0.0 : grad_out[keep_excl[i]];   // SparseScatterUpdate.h:52-57
}

struct SsuPlan {
    DevBuf<uint8_t> rmask, cmask;
    DevBuf<uint32_t> keep;
    uint32_t kept = 0;
};
static int ssu_plan(SsuPlan& P, int64_t on, const int64_t* oii, const int64_t* ojj, int64_t m, int64_t n,
                    const int64_t* ii, int64_t ni, const int64_t* jj, int64_t nj) {
    DevBuf<int> err; DevBuf<uint32_t> total;
    AB_TRY(err.alloc(1)); AB_TRY(total.alloc(1));
    AB_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), stream()));
    AB_TRY(P.rmask.alloc((size_t)m)); AB_TRY(P.cmask.alloc((size_t)n)); AB_TRY(P.keep.alloc((size_t)on));
    AB_CUDA(cudaMemsetAsync(P.rmask.p, 0, (size_t)(m > 0 ? m : 1), stream()));
    AB_CUDA(cudaMemsetAsync(P.cmask.p, 0, (size_t)(n > 0 ? n : 1), stream()));
    if (ni) { mark_mask_kernel<<<nblk(ni), 256, 0, stream()>>>(ii, (size_t)ni, 1, m, P.rmask.p, err.p); AB_LAUNCH_CHECK(); }
    if (nj) { mark_mask_kernel<<<nblk(nj), 256, 0, stream()>>>(jj, (size_t)nj, 1, n, P.cmask.p, err.p); AB_LAUNCH_CHECK(); }
    if (on) {
        ssu_flags_kernel<<<nblk(on), 256, 0, stream()>>>(oii, ojj, (size_t)on, P.rmask.p, P.cmask.p, m, n, P.keep.p, err.p);
        AB_LAUNCH_CHECK();
    }
    AB_TRY(scan_exclusive_u32(P.keep.p, P.keep.p, (size_t)on, total.p));
    int herr = 0;
    AB_CUDA(cudaMemcpyAsync(&P.kept, total.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream()));
    AB_TRY(fetch_err(err.p, &herr));
    if (herr) return fail(AB_ERANGE, "sparse_scatter_update: index outside the %lld x %lld matrix", (long long)m, (long long)n);
    return AB_OK;
}

// ================================================================ sparse_indexing
__global__ void sidx_flags_kernel(const int32_t* __restrict__ entry_row, const int32_t* __restrict__ colind, size_t nnz,
                                  const int32_t* __restrict__ rmap, const int32_t* __restrict__ cmap,
                                  uint32_t* __restrict__ keep) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    keep[e] = (rmap[entry_row[e]] >= 0 && cmap[colind[e]] >= 0) ? 1u : 0u;
}
// key = (position of the column in jj, row value): the reference walks jj in order and, inside a
// column, the intersection of the sorted row lists in ascending row order (SparseIndexing.h:44-61)
__global__ void sidx_keys_kernel(const int32_t* __restrict__ entry_row, const int32_t* __restrict__ colind, size_t nnz,
                                 const int32_t* __restrict__ rmap, const int32_t* __restrict__ cmap,
                                 const uint32_t* __restrict__ excl, int rowbits, uint64_t* __restrict__ keys,
                                 uint32_t* __restrict__ ent) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    const int r = entry_row[e], c = colind[e];
    if (rmap[r] < 0 || cmap[c] < 0) return;
    const uint32_t p = excl[e];
    keys[p] = ((uint64_t)(uint32_t)cmap[c] << rowbits) | (uint64_t)(uint32_t)r;
    ent[p]  = (uint32_t)e;
}
__global__ void sidx_emit_kernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ ent, size_t nout,
                                 const int32_t* __restrict__ entry_row, const int32_t* __restrict__ colind,
                                 const double* __restrict__ values, const int32_t* __restrict__ rmap,
                                 const int32_t* __restrict__ cmap, int64_t* __restrict__ oi, int64_t* __restrict__ oj,
                                 double* __restrict__ ov) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nout) return;
    const uint32_t e = ent[order[k]];
    oi[k] = (int64_t)rmap[entry_row[e]] + 1;
    oj[k] = (int64_t)cmap[colind[e]] + 1;
    ov[k] = values[e];
}
// backward (SparseIndexing.h:66-83): output i sits at (ii[ii0-1], jj[jj0-1]) of A; its gradient goes
// to the LAST input triplet carrying that (row, col) — std::map<pair,int> keeps the last index.
__global__ void sidx_grad_kernel(const int64_t* __restrict__ ii0, const int64_t* __restrict__ jj0,
                                 const double* __restrict__ g0, size_t nout, const int64_t* __restrict__ ii, int64_t ni,
                                 const int64_t* __restrict__ jj, int64_t nj, const int32_t* __restrict__ rowptr,
                                 const int32_t* __restrict__ colind, const uint32_t* __restrict__ segstart,
                                 const uint32_t* __restrict__ perm, int64_t m, int64_t n, uint64_t* __restrict__ tkey,
                                 int* __restrict__ err) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nout) return;
    tkey[k] = ~0ull;            // error exits leave a key no target carries
    const int64_t a = ii0[k] - 1, b = jj0[k] - 1;
    if (a < 0 || a >= ni || b < 0 || b >= nj) { atomicExch(err, 1); return; }
    const int64_t r = ii[a] - 1, c = jj[b] - 1;
    if (r < 0 || r >= m || c < 0 || c >= n) { atomicExch(err, 1); return; }
    int lo = rowptr[r], hi = rowptr[r + 1] - 1, e = -1;
    while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const int cm = colind[mid];
        if (cm == c) { e = mid; break; }
        if (cm < c) lo = mid + 1; else hi = mid - 1;
    }
    if (e < 0) { atomicExch(err, 2); return; }
    const uint32_t t = perm ? perm[segstart[e + 1] - 1] : (uint32_t)e;
    tkey[k] = (uint64_t)t;      // summed per target in ascending k by sidx_grad_reduce_kernel (no fp atomics)
}
// keys sorted by target triplet (stable: ascending k inside a target); the head of each run adds its run left to
// right — the order of the reference's sequential loop — so repeated (ii, jj) positions give a reproducible sum
__global__ void sidx_grad_reduce_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ kidx, size_t nout,
                                        const double* __restrict__ g0, double* __restrict__ g1) {
    size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nout) return;
    const uint64_t key = keys[p];
    if (key == ~0ull || (p > 0 && keys[p - 1] == key)) return;
    double sum = g0[kidx[p]];
    for (size_t q = p + 1; q < nout && keys[q] == key; ++q) sum += g0[kidx[q]];
    g1[key] = sum;
}

// ================================================================ CSR -> COO, scaling, SpGEMM expansion
__global__ void csr_to_coo_kernel(const int32_t* __restrict__ entry_row, const int32_t* __restrict__ colind,
                                  const double* __restrict__ values, size_t nnz, int base, double scale,
                                  int64_t* __restrict__ ii, int64_t* __restrict__ jj, double* __restrict__ vv) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    if (ii) ii[e] = (int64_t)entry_row[e] + base;
    if (jj) jj[e] = (int64_t)colind[e] + base;
    if (vv) vv[e] = scale == 1.0 ? values[e] : scale * values[e];
}
__global__ void csr_to_coo32_kernel(const int32_t* __restrict__ entry_row, const int32_t* __restrict__ colind,
                                    const double* __restrict__ values, size_t nnz, double scale,
                                    int32_t* __restrict__ ii, int32_t* __restrict__ jj, double* __restrict__ vv) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    ii[e] = entry_row[e]; jj[e] = colind[e];
    vv[e] = scale == 1.0 ? values[e] : scale * values[e];
}
// out[t] = vv[t] * d[idx[t] - base]   (forward_diag_sparse / forward_sparse_diag, SparseMatMul.h:34-52)
__global__ void coo_scale_kernel(const int64_t* __restrict__ idx, const double* __restrict__ vv, size_t T,
                                 const double* __restrict__ d, int64_t nd, int64_t base, double* __restrict__ out,
                                 int* __restrict__ err) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int64_t k = idx[t] - base;
    if (k < 0 || k >= nd) { atomicExch(err, 1); out[t] = 0.0; return; }
    out[t] = vv[t] * d[k];
}
// [ii1; ii2 + di], [jj1; jj2 + dj], [vv1; vv2]   (SparseConcate.h:11-27)
__global__ void concat_kernel(const int64_t* __restrict__ i1, const int64_t* __restrict__ j1, const double* __restrict__ v1,
                              size_t n1, const int64_t* __restrict__ i2, const int64_t* __restrict__ j2,
                              const double* __restrict__ v2, size_t n2, int64_t di, int64_t dj,
                              int64_t* __restrict__ oi, int64_t* __restrict__ oj, double* __restrict__ ov) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n1) { oi[k] = i1[k]; oj[k] = j1[k]; ov[k] = v1[k]; }
    else if (k < n1 + n2) { const size_t q = k - n1; oi[k] = i2[q] + di; oj[k] = j2[q] + dj; ov[k] = v2[q]; }
}
// dense[row, col] = value for every stored entry (duplicates were summed at assembly; SparseToDense.h:1-6)
__global__ void csr_to_dense_kernel(const int32_t* __restrict__ entry_row, const int32_t* __restrict__ colind,
                                    const double* __restrict__ values, size_t nnz, int64_t n, double* __restrict__ out) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < nnz) out[(size_t)entry_row[e] * (size_t)n + (size_t)colind[e]] = values[e];
}
// per-input-triplet gradient of the dense form: grad_v[t] = grad_out[row_t, col_t]  (SparseToDense.h:8-12)
__global__ void dense_grad_kernel(const int64_t* __restrict__ idx2, size_t T, int64_t m, int64_t n,
                                  const double* __restrict__ grad_out, double* __restrict__ grad_v, int* __restrict__ err) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const int64_t r = idx2[2 * t], c = idx2[2 * t + 1];
    if (r < 0 || r >= m || c < 0 || c >= n) { atomicExch(err, 1); grad_v[t] = 0.0; return; }
    grad_v[t] = grad_out[(size_t)r * (size_t)n + (size_t)c];
}

// dense_to_sparse (src/sparse.jl:78-86: tf.where(o != 0) + gather_nd): row-major nonzeros
__global__ void dense_flags_kernel(const double* __restrict__ a, size_t mn, uint32_t* __restrict__ keep) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < mn) keep[i] = (a[i] != 0.0) ? 1u : 0u;           // NaN != 0 is true, as in tf.not_equal
}
__global__ void dense_compact_kernel(const double* __restrict__ a, size_t mn, int64_t n, const uint32_t* __restrict__ excl,
                                     int64_t* __restrict__ oi, int64_t* __restrict__ oj, double* __restrict__ ov) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= mn) return;
    const double v = a[i];
    if (!(v != 0.0)) return;
    const uint32_t p = excl[i];
    oi[p] = (int64_t)(i / (size_t)n) + 1; oj[p] = (int64_t)(i % (size_t)n) + 1; ov[p] = v;
}

// number of products each stored entry of A generates = length of the matching row of B
__global__ void spgemm_count_kernel(const int32_t* __restrict__ a_col, size_t a_nnz, const int32_t* __restrict__ b_rowptr,
                                    uint32_t* __restrict__ cnt, unsigned long long* __restrict__ total) {
    __shared__ unsigned long long s_sum;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t c = 0;
    if (e < a_nnz) { const int k = a_col[e]; c = (uint32_t)(b_rowptr[k + 1] - b_rowptr[k]); cnt[e] = c; }
    unsigned long long w = c;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if ((threadIdx.x & 31) == 0 && w) atomicAdd(&s_sum, w);
    __syncthreads();
    if (threadIdx.x == 0 && s_sum) atomicAdd(total, s_sum);
}
// products in (row of A, k ascending, column of B ascending) order: for a fixed C(i,j) the stream
// order is k-ascending, which the stable sort-and-reduce of assemble.cu preserves.
// One warp per stored entry of A; lanes stride the row of B.
__global__ void __launch_bounds__(256)
spgemm_expand_kernel(const int32_t* __restrict__ a_row, const int32_t* __restrict__ a_col,
                     const double* __restrict__ a_val, size_t a_nnz, const uint32_t* __restrict__ off,
                     const int32_t* __restrict__ b_rowptr, const int32_t* __restrict__ b_col,
                     const double* __restrict__ b_val, int swap, int32_t* __restrict__ pi, int32_t* __restrict__ pj,
                     double* __restrict__ pv) {
    const int lane = threadIdx.x & 31;
    const size_t w0 = ((size_t)blockIdx.x * 256 + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * 256) >> 5;
    for (size_t e = w0; e < a_nnz; e += nw) {
        const int i = a_row[e], k = a_col[e];
        const double a = a_val[e];
        const int b0 = b_rowptr[k], b1 = b_rowptr[k + 1];
        const size_t o = off[e];
        for (int q = b0 + lane; q < b1; q += 32) {
            const size_t p = o + (size_t)(q - b0);
            const int j = b_col[q];
            pi[p] = swap ? j : i;
            pj[p] = swap ? i : j;
            pv[p] = __dmul_rn(a, b_val[q]);
        }
    }
}

}  // namespace ab

using namespace ab;

extern "C" {

// ---------------------------------------------------------------- triplet handles
int64_t ab_trip_size(int64_t th) {
    Trip* t = trip_get(th);
    return t ? t->n : (int64_t)AB_EHANDLE;
}
int ab_trip_destroy(int64_t th) {
    AB_TRY(ensure_device());
    Trip* t = trip_take(th);
    if (!t) return AB_EHANDLE;
    delete t;
    return AB_OK;
}
int ab_trip_fetch(int64_t th, int64_t* ii, int64_t* jj, double* vv) {
    AB_TRY(ensure_device());
    Trip* t = trip_take(th);
    if (!t) return AB_EHANDLE;
    std::lock_guard<std::mutex> lk(big_lock());
    int rc = AB_OK;
    auto body = [&]() -> int {
        if (t->n) {
            if (ii) AB_CUDA(cudaMemcpyAsync(ii, t->ii, (size_t)t->n * sizeof(int64_t), cudaMemcpyDefault, stream()));
            if (jj) AB_CUDA(cudaMemcpyAsync(jj, t->jj, (size_t)t->n * sizeof(int64_t), cudaMemcpyDefault, stream()));
            if (vv) AB_CUDA(cudaMemcpyAsync(vv, t->vv, (size_t)t->n * sizeof(double), cudaMemcpyDefault, stream()));
        }
        AB_CUDA(cudaStreamSynchronize(stream()));
        return AB_OK;
    };
    rc = body();
    delete t;
    return rc;
}
// interleaved [n,2] index matrix (the layout of tf.SparseTensor.indices, used by zero_out_row)
int ab_trip_fetch_idx2(int64_t th, int64_t* idx2, double* vv) {
    AB_TRY(ensure_device());
    Trip* t = trip_take(th);
    if (!t) return AB_EHANDLE;
    std::lock_guard<std::mutex> lk(big_lock());
    auto body = [&]() -> int {
        if (t->n) {
            if (idx2) {
                AB_CUDA(cudaMemcpy2DAsync(idx2, 2 * sizeof(int64_t), t->ii, sizeof(int64_t), sizeof(int64_t), (size_t)t->n,
                                          cudaMemcpyDefault, stream()));
                AB_CUDA(cudaMemcpy2DAsync(idx2 + 1, 2 * sizeof(int64_t), t->jj, sizeof(int64_t), sizeof(int64_t), (size_t)t->n,
                                          cudaMemcpyDefault, stream()));
            }
            if (vv) AB_CUDA(cudaMemcpyAsync(vv, t->vv, (size_t)t->n * sizeof(double), cudaMemcpyDefault, stream()));
        }
        AB_CUDA(cudaStreamSynchronize(stream()));
        return AB_OK;
    };
    int rc = body();
    delete t;
    return rc;
}
int ab_csr_from_trip(int64_t th, int64_t m, int64_t n, int base, int64_t* out_h) {
    AB_TRY(ensure_device());
    Trip* t = trip_take(th);
    if (!t) return AB_EHANDLE;
    int rc = ab_csr_from_coo(t->n, t->ii, t->jj, t->vv, m, n, base, out_h);
    std::lock_guard<std::mutex> lk(big_lock());
    delete t;
    return rc;
}

// ---------------------------------------------------------------- zero_out_row
int ab_zero_out_row(int64_t N, const int64_t* idx2_, const double* v_, const int64_t* bd_, int64_t nbd,
                    int64_t* out_th) {
    AB_TRY(ensure_device());
    if (!out_th) return fail(AB_EINVAL, "out_th is null");
    if (N < 0 || nbd < 0) return fail(AB_EINVAL, "negative size");
    if ((N > 0 && (!idx2_ || !v_)) || (nbd > 0 && !bd_)) return fail(AB_EINVAL, "null argument");
    if (N >= (int64_t)UINT32_MAX) return fail(AB_EUNSUPPORTED, "more than 2^32-1 entries in one call");
    std::lock_guard<std::mutex> lk(big_lock());
    In<int64_t> idx2, bd; In<double> v;
    AB_TRY(idx2.bind(idx2_, (size_t)N * 2)); AB_TRY(v.bind(v_, (size_t)N)); AB_TRY(bd.bind(bd_, (size_t)nbd));
    DevBuf<uint8_t> mask; int64_t masklen = 0;
    AB_TRY(zor_mask(bd.d, nbd, mask, &masklen));
    DevBuf<uint32_t> keep, total;
    AB_TRY(keep.alloc((size_t)N)); AB_TRY(total.alloc(1));
    if (N) { zor_flags_kernel<<<nblk(N), 256, 0, stream()>>>(idx2.d, (size_t)N, mask.p, masklen, keep.p); AB_LAUNCH_CHECK(); }
    AB_TRY(scan_exclusive_u32(keep.p, keep.p, (size_t)N, total.p));
    uint32_t nout = 0;
    AB_TRY(fetch_u32(total.p, &nout));
    Trip* t = new Trip();
    auto body = [&]() -> int {
        AB_TRY(trip_alloc(t, nout));
        if (N) {
            zor_compact_kernel<<<nblk(N), 256, 0, stream()>>>(idx2.d, v.d, (size_t)N, mask.p, masklen, keep.p, t->ii, t->jj, t->vv);
            AB_LAUNCH_CHECK();
        }
        AB_CUDA(cudaStreamSynchronize(stream()));   // staged inputs die with this scope
        return AB_OK;
    };
    int rc = body();
    if (rc != AB_OK) { delete t; return rc; }
    *out_th = trip_put(t);
    return AB_OK;
}

int ab_zero_out_row_grad(int64_t N, const int64_t* idx2_, const int64_t* bd_, int64_t nbd, const double* grad_ov_,
                         int64_t nout, double* grad_v_) {
    AB_TRY(ensure_device());
    if (N < 0 || nbd < 0 || nout < 0) return fail(AB_EINVAL, "negative size");
    if (N == 0) return AB_OK;
    if (!idx2_ || !grad_v_ || (nbd > 0 && !bd_) || (nout > 0 && !grad_ov_)) return fail(AB_EINVAL, "null argument");
    if (N >= (int64_t)UINT32_MAX) return fail(AB_EUNSUPPORTED, "more than 2^32-1 entries in one call");
    std::lock_guard<std::mutex> lk(big_lock());
    In<int64_t> idx2, bd; In<double> gov; Out<double> gv;
    AB_TRY(idx2.bind(idx2_, (size_t)N * 2)); AB_TRY(bd.bind(bd_, (size_t)nbd));
    AB_TRY(gov.bind(grad_ov_, (size_t)nout)); AB_TRY(gv.bind(grad_v_, (size_t)N));
    DevBuf<uint8_t> mask; int64_t masklen = 0;
    AB_TRY(zor_mask(bd.d, nbd, mask, &masklen));
    DevBuf<uint32_t> keep, total;
    AB_TRY(keep.alloc((size_t)N)); AB_TRY(total.alloc(1));
    zor_flags_kernel<<<nblk(N), 256, 0, stream()>>>(idx2.d, (size_t)N, mask.p, masklen, keep.p);
    AB_LAUNCH_CHECK();
    AB_TRY(scan_exclusive_u32(keep.p, keep.p, (size_t)N, total.p));
    uint32_t kept = 0;
    AB_TRY(fetch_u32(total.p, &kept));
    if ((int64_t)kept != nout) return fail(AB_EINVAL, "zero_out_row_grad: grad_oval has %lld entries, forward kept %u", (long long)nout, kept);
    zor_grad_kernel<<<nblk(N), 256, 0, stream()>>>(idx2.d, (size_t)N, mask.p, masklen, keep.p, gov.d, gv.d);
    AB_LAUNCH_CHECK();
    AB_TRY(gv.commit());
    AB_CUDA(cudaStreamSynchronize(stream()));
    return AB_OK;
}

// ---------------------------------------------------------------- sparse_scatter_update
int ab_scatter_update(int64_t on, const int64_t* oii_, const int64_t* ojj_, const double* ovv_,
                      int64_t un, const int64_t* uii_, const int64_t* ujj_, const double* uvv_,
                      int64_t m, int64_t n, const int64_t* ii_, int64_t ni, const int64_t* jj_, int64_t nj,
                      int64_t* out_th) {
    AB_TRY(ensure_device());
    if (!out_th) return fail(AB_EINVAL, "out_th is null");
    if (on < 0 || un < 0 || m < 0 || n < 0 || ni < 0 || nj < 0) return fail(AB_EINVAL, "negative size");
    if ((on > 0 && (!oii_ || !ojj_ || !ovv_)) || (un > 0 && (!uii_ || !ujj_ || !uvv_)) || (ni > 0 && !ii_) || (nj > 0 && !jj_))
        return fail(AB_EINVAL, "null argument");
    if (on + un >= (int64_t)UINT32_MAX) return fail(AB_EUNSUPPORTED, "more than 2^32-1 entries in one call");
    std::lock_guard<std::mutex> lk(big_lock());
    In<int64_t> oii, ojj, uii, ujj, ii, jj; In<double> ovv, uvv;
    AB_TRY(oii.bind(oii_, on)); AB_TRY(ojj.bind(ojj_, on)); AB_TRY(ovv.bind(ovv_, on));
    AB_TRY(uii.bind(uii_, un)); AB_TRY(ujj.bind(ujj_, un)); AB_TRY(uvv.bind(uvv_, un));
    AB_TRY(ii.bind(ii_, ni)); AB_TRY(jj.bind(jj_, nj));
    SsuPlan P;
    AB_TRY(ssu_plan(P, on, oii.d, ojj.d, m, n, ii.d, ni, jj.d, nj));
    Trip* t = new Trip();
    auto body = [&]() -> int {
        AB_TRY(trip_alloc(t, (int64_t)P.kept + un));
        DevBuf<int> err;
        AB_TRY(err.alloc(1));
        AB_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), stream()));
        if (on) {
            ssu_compact_kernel<<<nblk(on), 256, 0, stream()>>>(oii.d, ojj.d, ovv.d, (size_t)on, P.keep.p, P.rmask.p, P.cmask.p,
                                                               t->ii, t->jj, t->vv);
            AB_LAUNCH_CHECK();
        }
        if (un) {
            ssu_append_kernel<<<nblk(un), 256, 0, stream()>>>(uii.d, ujj.d, uvv.d, (size_t)un, ii.d, ni, jj.d, nj,
                                                              t->ii + P.kept, t->jj + P.kept, t->vv + P.kept, err.p);
            AB_LAUNCH_CHECK();
        }
        int herr = 0;
        AB_TRY(fetch_err(err.p, &herr));
        if (herr) return fail(AB_ERANGE, "sparse_scatter_update: block index outside [1,%lld]x[1,%lld]", (long long)ni, (long long)nj);
        return AB_OK;
    };
    int rc = body();
    if (rc != AB_OK) { delete t; return rc; }
    *out_th = trip_put(t);
    return AB_OK;
}

int ab_scatter_update_grad(int64_t on, const int64_t* oii_, const int64_t* ojj_, int64_t un, int64_t m, int64_t n,
                           const int64_t* ii_, int64_t ni, const int64_t* jj_, int64_t nj, const double* grad_out_,
                           int64_t nout, double* grad_ovv_, double* grad_uvv_) {
    AB_TRY(ensure_device());
    if (on < 0 || un < 0 || m < 0 || n < 0 || ni < 0 || nj < 0 || nout < 0) return fail(AB_EINVAL, "negative size");
    if ((on > 0 && (!oii_ || !ojj_)) || (ni > 0 && !ii_) || (nj > 0 && !jj_) || (nout > 0 && !grad_out_))
        return fail(AB_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(big_lock());
    In<int64_t> oii, ojj, ii, jj; In<double> go; Out<double> govv, guvv;
    AB_TRY(oii.bind(oii_, on)); AB_TRY(ojj.bind(ojj_, on)); AB_TRY(ii.bind(ii_, ni)); AB_TRY(jj.bind(jj_, nj));
    AB_TRY(go.bind(grad_out_, nout)); AB_TRY(govv.bind(grad_ovv_, on)); AB_TRY(guvv.bind(grad_uvv_, un));
    SsuPlan P;
    AB_TRY(ssu_plan(P, on, oii.d, ojj.d, m, n, ii.d, ni, jj.d, nj));
    if ((int64_t)P.kept + un != nout)
        return fail(AB_EINVAL, "scatter_update_grad: grad_out has %lld entries, forward produced %lld", (long long)nout,
                    (long long)(P.kept + un));
    if (on && govv.d) {
        ssu_grad_kernel<<<nblk(on), 256, 0, stream()>>>(oii.d, ojj.d, (size_t)on, P.keep.p, P.rmask.p, P.cmask.p, go.d, govv.d);
        AB_LAUNCH_CHECK();
    }
    if (un && guvv.d)   // SparseScatterUpdate.h:58-60
        AB_CUDA(cudaMemcpyAsync(guvv.d, go.d + P.kept, (size_t)un * sizeof(double), cudaMemcpyDeviceToDevice, stream()));
    AB_TRY(govv.commit()); AB_TRY(guvv.commit());
    AB_CUDA(cudaStreamSynchronize(stream()));
    return AB_OK;
}

// ---------------------------------------------------------------- sparse_indexing  A[ii, jj]
int ab_sparse_indexing(int64_t h, const int64_t* ii_, int64_t ni, const int64_t* jj_, int64_t nj, int64_t* out_th) {
    AB_TRY(ensure_device());
    Csr* c = csr_get(h);
    if (!c) return AB_EHANDLE;
    if (!out_th) return fail(AB_EINVAL, "out_th is null");
    if (ni < 0 || nj < 0 || (ni > 0 && !ii_) || (nj > 0 && !jj_)) return fail(AB_EINVAL, "bad index list");
    if (ni >= INT32_MAX || nj >= INT32_MAX) return fail(AB_EUNSUPPORTED, "index list exceeds int32");
    std::lock_guard<std::mutex> lk(big_lock());
    In<int64_t> ii, jj;
    AB_TRY(ii.bind(ii_, ni)); AB_TRY(jj.bind(jj_, nj));
    AB_TRY(csr_ensure_entry_row(c));
    const size_t nnz = (size_t)c->nnz;
    DevBuf<int32_t> rmap, cmap; DevBuf<int> err; DevBuf<uint32_t> keep, total;
    AB_TRY(rmap.alloc((size_t)c->m)); AB_TRY(cmap.alloc((size_t)c->n)); AB_TRY(err.alloc(1));
    AB_TRY(keep.alloc(nnz)); AB_TRY(total.alloc(1));
    AB_CUDA(cudaMemsetAsync(rmap.p, 0xff, (size_t)(c->m > 0 ? c->m : 1) * sizeof(int32_t), stream()));
    AB_CUDA(cudaMemsetAsync(cmap.p, 0xff, (size_t)(c->n > 0 ? c->n : 1) * sizeof(int32_t), stream()));
    AB_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), stream()));
    if (ni) { mark_map_kernel<<<nblk(ni), 256, 0, stream()>>>(ii.d, (size_t)ni, c->m, rmap.p, err.p); AB_LAUNCH_CHECK(); }
    if (nj) { mark_map_kernel<<<nblk(nj), 256, 0, stream()>>>(jj.d, (size_t)nj, c->n, cmap.p, err.p); AB_LAUNCH_CHECK(); }
    if (nnz) { sidx_flags_kernel<<<nblk(nnz), 256, 0, stream()>>>(c->entry_row, c->colind, nnz, rmap.p, cmap.p, keep.p); AB_LAUNCH_CHECK(); }
    AB_TRY(scan_exclusive_u32(keep.p, keep.p, nnz, total.p));
    uint32_t nout = 0; int herr = 0;
    AB_CUDA(cudaMemcpyAsync(&nout, total.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream()));
    AB_TRY(fetch_err(err.p, &herr));
    if (herr) return fail(AB_ERANGE, "sparse_indexing: index outside the %lld x %lld matrix", (long long)c->m, (long long)c->n);
    Trip* t = new Trip();
    auto body = [&]() -> int {
        AB_TRY(trip_alloc(t, nout));
        if (nout == 0) return AB_OK;
        DevBuf<uint64_t> k0, k1; DevBuf<uint32_t> p0, p1, ent;
        AB_TRY(k0.alloc(nout)); AB_TRY(k1.alloc(nout)); AB_TRY(p0.alloc(nout)); AB_TRY(p1.alloc(nout)); AB_TRY(ent.alloc(nout));
        const int rowbits = bits_for(c->m > 0 ? c->m : 1);
        sidx_keys_kernel<<<nblk(nnz), 256, 0, stream()>>>(c->entry_row, c->colind, nnz, rmap.p, cmap.p, keep.p, rowbits, k0.p, ent.p);
        AB_LAUNCH_CHECK();
        uint64_t* ks; uint32_t* ps;
        AB_TRY(radix_sort_u64(k0.p, p0.p, k1.p, p1.p, nout, rowbits + bits_for(nj > 0 ? nj : 1), &ks, &ps));
        sidx_emit_kernel<<<nblk(nout), 256, 0, stream()>>>(ps, ent.p, nout, c->entry_row, c->colind, c->values, rmap.p, cmap.p,
                                                          t->ii, t->jj, t->vv);
        AB_LAUNCH_CHECK();
        AB_CUDA(cudaStreamSynchronize(stream()));
        return AB_OK;
    };
    int rc = body();
    if (rc != AB_OK) { delete t; return rc; }
    *out_th = trip_put(t);
    return AB_OK;
}

int ab_sparse_indexing_grad(int64_t h, const int64_t* ii_, int64_t ni, const int64_t* jj_, int64_t nj,
                            const int64_t* ii0_, const int64_t* jj0_, const double* grad_vv0_, int64_t nout,
                            double* grad_vv1_) {
    AB_TRY(ensure_device());
    Csr* c = csr_get(h);
    if (!c) return AB_EHANDLE;
    if (ni < 0 || nj < 0 || nout < 0) return fail(AB_EINVAL, "negative size");
    if ((ni > 0 && !ii_) || (nj > 0 && !jj_) || (nout > 0 && (!ii0_ || !jj0_ || !grad_vv0_)) || (c->T > 0 && !grad_vv1_))
        return fail(AB_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(big_lock());
    In<int64_t> ii, jj, ii0, jj0; In<double> g0; Out<double> g1;
    AB_TRY(ii.bind(ii_, ni)); AB_TRY(jj.bind(jj_, nj)); AB_TRY(ii0.bind(ii0_, nout)); AB_TRY(jj0.bind(jj0_, nout));
    AB_TRY(g0.bind(grad_vv0_, nout)); AB_TRY(g1.bind(grad_vv1_, (size_t)c->T));
    if (c->T) AB_CUDA(cudaMemsetAsync(g1.d, 0, (size_t)c->T * sizeof(double), stream()));
    DevBuf<int> err;
    AB_TRY(err.alloc(1));
    AB_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), stream()));
    if (nout) {
        DevBuf<uint64_t> k0, k1; DevBuf<uint32_t> p0, p1;
        AB_TRY(k0.alloc(nout)); AB_TRY(k1.alloc(nout)); AB_TRY(p0.alloc(nout)); AB_TRY(p1.alloc(nout));
        sidx_grad_kernel<<<nblk(nout), 256, 0, stream()>>>(ii0.d, jj0.d, g0.d, (size_t)nout, ii.d, ni, jj.d, nj, c->rowptr, c->colind,
                                                          c->segstart, c->perm, c->m, c->n, k0.p, err.p);
        AB_LAUNCH_CHECK();
        uint64_t* ks; uint32_t* ps;
        int nbits = 1;
        while (nbits < 40 && ((int64_t)1 << nbits) < c->T) ++nbits;
        AB_TRY(radix_sort_u64(k0.p, p0.p, k1.p, p1.p, (size_t)nout, nbits, &ks, &ps));  // error keys (~0) abort the call below anyway
        sidx_grad_reduce_kernel<<<nblk(nout), 256, 0, stream()>>>(ks, ps, (size_t)nout, g0.d, g1.d);
        AB_LAUNCH_CHECK();
    }
    int herr = 0;
    AB_TRY(fetch_err(err.p, &herr));
    if (herr == 1) return fail(AB_ERANGE, "sparse_indexing_grad: index out of range");
    if (herr == 2) return fail(AB_EINVAL, "sparse_indexing_grad: an output position is not a stored entry of A");
    AB_TRY(g1.commit());
    AB_CUDA(cudaStreamSynchronize(stream()));
    return AB_OK;
}

// ---------------------------------------------------------------- CSR -> COO export, diag scaling
int ab_csr_export_coo(int64_t h, int64_t* ii_, int64_t* jj_, double* vv_, int base) {
    AB_TRY(ensure_device());
    Csr* c = csr_get(h);
    if (!c) return AB_EHANDLE;
    if (c->nnz == 0) return AB_OK;
    std::lock_guard<std::mutex> lk(big_lock());
    AB_TRY(csr_ensure_entry_row(c));
    Out<int64_t> ii, jj; Out<double> vv;
    AB_TRY(ii.bind(ii_, (size_t)c->nnz)); AB_TRY(jj.bind(jj_, (size_t)c->nnz)); AB_TRY(vv.bind(vv_, (size_t)c->nnz));
    csr_to_coo_kernel<<<nblk(c->nnz), 256, 0, stream()>>>(c->entry_row, c->colind, c->values, (size_t)c->nnz, base, 1.0, ii.d, jj.d, vv.d);
    AB_LAUNCH_CHECK();
    AB_TRY(ii.commit()); AB_TRY(jj.commit()); AB_TRY(vv.commit());
    AB_CUDA(cudaStreamSynchronize(stream()));
    return AB_OK;
}

int ab_coo_scale(int64_t T, const int64_t* idx_, const double* vv_, const double* d_, int64_t nd, int base, double* out_) {
    AB_TRY(ensure_device());
    if (T < 0 || nd < 0) return fail(AB_EINVAL, "negative size");
    if (T == 0) return AB_OK;
    if (!idx_ || !vv_ || !d_ || !out_) return fail(AB_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(big_lock());
    In<int64_t> idx; In<double> vv, d; Out<double> out;
    AB_TRY(idx.bind(idx_, T)); AB_TRY(vv.bind(vv_, T)); AB_TRY(d.bind(d_, nd)); AB_TRY(out.bind(out_, T));
    DevBuf<int> err;
    AB_TRY(err.alloc(1));
    AB_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), stream()));
    coo_scale_kernel<<<nblk(T), 256, 0, stream()>>>(idx.d, vv.d, (size_t)T, d.d, nd, base, out.d, err.p);
    AB_LAUNCH_CHECK();
    int herr = 0;
    AB_TRY(fetch_err(err.p, &herr));
    if (herr) return fail(AB_ERANGE, "coo_scale: index outside the diagonal of length %lld", (long long)nd);
    AB_TRY(out.commit());
    AB_CUDA(cudaStreamSynchronize(stream()));
    return AB_OK;
}

// ---------------------------------------------------------------- dense_to_sparse
int ab_dense_to_coo(int64_t m, int64_t n, const double* a_, int64_t* out_th) {
    AB_TRY(ensure_device());
    if (!out_th) return fail(AB_EINVAL, "out_th is null");
    if (m < 0 || n < 0) return fail(AB_EINVAL, "negative size");
    const size_t mn = (size_t)m * (size_t)n;
    if (mn > 0 && !a_) return fail(AB_EINVAL, "null matrix");
    if (mn >= (size_t)UINT32_MAX) return fail(AB_EUNSUPPORTED, "more than 2^32-1 dense entries in one call");
    std::lock_guard<std::mutex> lk(big_lock());
    In<double> a;
    AB_TRY(a.bind(a_, mn));
    DevBuf<uint32_t> keep, total;
    AB_TRY(keep.alloc(mn)); AB_TRY(total.alloc(1));
    if (mn) { dense_flags_kernel<<<nblk(mn), 256, 0, stream()>>>(a.d, mn, keep.p); AB_LAUNCH_CHECK(); }
    AB_TRY(scan_exclusive_u32(keep.p, keep.p, mn, total.p));
    uint32_t nout = 0;
    AB_TRY(fetch_u32(total.p, &nout));
    Trip* t = new Trip();
    auto body = [&]() -> int {
        AB_TRY(trip_alloc(t, nout));
        if (mn) { dense_compact_kernel<<<nblk(mn), 256, 0, stream()>>>(a.d, mn, n, keep.p, t->ii, t->jj, t->vv); AB_LAUNCH_CHECK(); }
        AB_CUDA(cudaStreamSynchronize(stream()));
        return AB_OK;
    };
    int rc = body();
    if (rc != AB_OK) { delete t; return rc; }
    *out_th = trip_put(t);
    return AB_OK;
}

// ---------------------------------------------------------------- vcat / hcat (sparse_concate)
int ab_coo_concat(int64_t n1, const int64_t* ii1_, const int64_t* jj1_, const double* vv1_,
                  int64_t n2, const int64_t* ii2_, const int64_t* jj2_, const double* vv2_,
                  int64_t m1, int64_t ncol1, int hcat, int64_t* oii_, int64_t* ojj_, double* ovv_) {
    AB_TRY(ensure_device());
    if (n1 < 0 || n2 < 0) return fail(AB_EINVAL, "negative size");
    if (n1 + n2 == 0) return AB_OK;
    if ((n1 > 0 && (!ii1_ || !jj1_ || !vv1_)) || (n2 > 0 && (!ii2_ || !jj2_ || !vv2_)) || !oii_ || !ojj_ || !ovv_)
        return fail(AB_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(big_lock());
    In<int64_t> i1, j1, i2, j2; In<double> v1, v2; Out<int64_t> oi, oj; Out<double> ov;
    AB_TRY(i1.bind(ii1_, n1)); AB_TRY(j1.bind(jj1_, n1)); AB_TRY(v1.bind(vv1_, n1));
    AB_TRY(i2.bind(ii2_, n2)); AB_TRY(j2.bind(jj2_, n2)); AB_TRY(v2.bind(vv2_, n2));
    AB_TRY(oi.bind(oii_, n1 + n2)); AB_TRY(oj.bind(ojj_, n1 + n2)); AB_TRY(ov.bind(ovv_, n1 + n2));
    concat_kernel<<<nblk(n1 + n2), 256, 0, stream()>>>(i1.d, j1.d, v1.d, (size_t)n1, i2.d, j2.d, v2.d, (size_t)n2,
                                                       hcat ? 0 : m1, hcat ? ncol1 : 0, oi.d, oj.d, ov.d);
    AB_LAUNCH_CHECK();
    AB_TRY(oi.commit()); AB_TRY(oj.commit()); AB_TRY(ov.commit());
    AB_CUDA(cudaStreamSynchronize(stream()));
    return AB_OK;
}

// ---------------------------------------------------------------- Array(A) (sparse_to_dense_ad)
int ab_csr_to_dense(int64_t h, double* out_) {
    AB_TRY(ensure_device());
    Csr* c = csr_get(h);
    if (!c) return AB_EHANDLE;
    if (c->m * c->n == 0) return AB_OK;
    if (!out_) return fail(AB_EINVAL, "out is null");
    std::lock_guard<std::mutex> lk(big_lock());
    AB_TRY(csr_ensure_entry_row(c));
    Out<double> out;
    AB_TRY(out.bind(out_, (size_t)(c->m * c->n)));
    AB_CUDA(cudaMemsetAsync(out.d, 0, (size_t)(c->m * c->n) * sizeof(double), stream()));
    if (c->nnz) {
        csr_to_dense_kernel<<<nblk(c->nnz), 256, 0, stream()>>>(c->entry_row, c->colind, c->values, (size_t)c->nnz, c->n, out.d);
        AB_LAUNCH_CHECK();
    }
    AB_TRY(out.commit());
    AB_CUDA(cudaStreamSynchronize(stream()));
    return AB_OK;
}
int ab_dense_grad_values(int64_t T, const int64_t* idx2_, int64_t m, int64_t n, const double* grad_out_, double* grad_v_) {
    AB_TRY(ensure_device());
    if (T < 0 || m < 0 || n < 0) return fail(AB_EINVAL, "negative size");
    if (T == 0) return AB_OK;
    if (!idx2_ || !grad_out_ || !grad_v_) return fail(AB_EINVAL, "null argument");
    std::lock_guard<std::mutex> lk(big_lock());
    In<int64_t> idx2; In<double> go; Out<double> gv;
    AB_TRY(idx2.bind(idx2_, (size_t)T * 2)); AB_TRY(go.bind(grad_out_, (size_t)(m * n))); AB_TRY(gv.bind(grad_v_, (size_t)T));
    DevBuf<int> err;
    AB_TRY(err.alloc(1));
    AB_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), stream()));
    dense_grad_kernel<<<nblk(T), 256, 0, stream()>>>(idx2.d, (size_t)T, m, n, go.d, gv.d, err.p);
    AB_LAUNCH_CHECK();
    int herr = 0;
    AB_TRY(fetch_err(err.p, &herr));
    if (herr) return fail(AB_ERANGE, "sparse_to_dense_grad: index outside the %lld x %lld matrix", (long long)m, (long long)n);
    AB_TRY(gv.commit());
    AB_CUDA(cudaStreamSynchronize(stream()));
    return AB_OK;
}

// ---------------------------------------------------------------- A + B (tf.sparse.add), C = alpha A + beta B
int ab_csr_axpby(int64_t ha, double alpha, int64_t hb, double beta, int64_t* out_h) {
    AB_TRY(ensure_device());
    Csr* a = csr_get(ha);
    if (!a) return AB_EHANDLE;
    Csr* b = csr_get(hb);
    if (!b) return AB_EHANDLE;
    if (!out_h) return fail(AB_EINVAL, "out_h is null");
    if (a->m != b->m || a->n != b->n)
        return fail(AB_EINVAL, "size (%lld, %lld) and (%lld, %lld) does not match", (long long)a->m, (long long)a->n,
                    (long long)b->m, (long long)b->n);
    const size_t na = (size_t)a->nnz, nb = (size_t)b->nnz;
    if (na + nb >= (size_t)UINT32_MAX) return fail(AB_EUNSUPPORTED, "more than 2^32-1 entries in one call");
    DevBuf<int32_t> pi, pj; DevBuf<double> pv;
    {
        std::lock_guard<std::mutex> lk(big_lock());
        AB_TRY(csr_ensure_entry_row(a)); AB_TRY(csr_ensure_entry_row(b));
        AB_TRY(pi.alloc(na + nb)); AB_TRY(pj.alloc(na + nb)); AB_TRY(pv.alloc(na + nb));
        if (na) { csr_to_coo32_kernel<<<nblk(na), 256, 0, stream()>>>(a->entry_row, a->colind, a->values, na, alpha, pi.p, pj.p, pv.p); AB_LAUNCH_CHECK(); }
        if (nb) { csr_to_coo32_kernel<<<nblk(nb), 256, 0, stream()>>>(b->entry_row, b->colind, b->values, nb, beta, pi.p + na, pj.p + na, pv.p + na); AB_LAUNCH_CHECK(); }
    }
    int rc = ab_csr_from_coo32((int64_t)(na + nb), pi.p, pj.p, pv.p, a->m, a->n, 0, out_h);
    std::lock_guard<std::mutex> lk(big_lock());
    pi.release(); pj.release(); pv.release();
    return rc;
}

// ---------------------------------------------------------------- SpGEMM  C = A B  (or C' when transpose_out)
int ab_spgemm(int64_t ha, int64_t hb, int transpose_out, int64_t* out_h) {
    AB_TRY(ensure_device());
    Csr* a = csr_get(ha);
    if (!a) return AB_EHANDLE;
    Csr* b = csr_get(hb);
    if (!b) return AB_EHANDLE;
    if (!out_h) return fail(AB_EINVAL, "out_h is null");
    if (a->n != b->m)
        return fail(AB_EINVAL, "matrix size mismatch: (%lld, %lld) vs (%lld, %lld)", (long long)a->m, (long long)a->n,
                    (long long)b->m, (long long)b->n);
    const size_t na = (size_t)a->nnz;
    DevBuf<int32_t> pi, pj; DevBuf<double> pv;
    unsigned long long nprod = 0;
    {
        std::lock_guard<std::mutex> lk(big_lock());
        AB_TRY(csr_ensure_entry_row(a));
        DevBuf<uint32_t> cnt; DevBuf<unsigned long long> total;
        AB_TRY(cnt.alloc(na)); AB_TRY(total.alloc(1));
        AB_CUDA(cudaMemsetAsync(total.p, 0, sizeof(unsigned long long), stream()));
        if (na) { spgemm_count_kernel<<<nblk(na), 256, 0, stream()>>>(a->colind, na, b->rowptr, cnt.p, total.p); AB_LAUNCH_CHECK(); }
        AB_CUDA(cudaMemcpyAsync(&nprod, total.p, sizeof(nprod), cudaMemcpyDeviceToHost, stream()));
        AB_CUDA(cudaStreamSynchronize(stream()));
        if (nprod >= (unsigned long long)UINT32_MAX)
            return fail(AB_EUNSUPPORTED, "SpGEMM: %llu intermediate products exceed the 2^32-1 per-call limit", nprod);
        AB_TRY(scan_exclusive_u32(cnt.p, cnt.p, na, nullptr));
        AB_TRY(pi.alloc((size_t)nprod)); AB_TRY(pj.alloc((size_t)nprod)); AB_TRY(pv.alloc((size_t)nprod));
        if (nprod) {
            size_t nb8 = (na + 7) / 8, cap = (size_t)sm_count() * 32;
            unsigned grid = (unsigned)(nb8 < cap ? nb8 : cap);
            spgemm_expand_kernel<<<grid, 256, 0, stream()>>>(a->entry_row, a->colind, a->values, na, cnt.p, b->rowptr, b->colind,
                                                            b->values, transpose_out ? 1 : 0, pi.p, pj.p, pv.p);
            AB_LAUNCH_CHECK();
        }
        AB_CUDA(cudaStreamSynchronize(stream()));
    }
    const int64_t cm = transpose_out ? b->n : a->m, cn = transpose_out ? a->m : b->n;
    int rc = ab_csr_from_coo32((int64_t)nprod, pi.p, pj.p, pv.p, cm, cn, 0, out_h);
    std::lock_guard<std::mutex> lk(big_lock());
    pi.release(); pj.release(); pv.release();
    return rc;
}

}  // extern "C"
