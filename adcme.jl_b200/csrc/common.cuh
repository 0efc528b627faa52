// common.cuh — shared host/device plumbing for libadcme_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <string.h>
#include <stdlib.h>
#include <mutex>
#include <vector>
#include "../../include/adcme_b200.h"

namespace ab {

// ---------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
int  fail(int code, const char* fmt, ...);

#define AB_CUDA(call)                                                                   \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess)                                                         \
            return ab::fail(AB_ECUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,      \
                            cudaGetErrorString(e__));                                   \
    } while (0)

#define AB_TRY(call)                       \
    do {                                   \
        int rc__ = (call);                 \
        if (rc__ != AB_OK) return rc__;    \
    } while (0)

#define AB_LAUNCH_CHECK()                                                               \
    do {                                                                                \
        ab::count_launch();                                                             \
        cudaError_t e__ = cudaPeekAtLastError();                                        \
        if (e__ != cudaSuccess)                                                         \
            return ab::fail(AB_ECUDA, "%s:%d kernel launch -> %s", __FILE__, __LINE__,  \
                            cudaGetErrorString(e__));                                   \
    } while (0)

// ---------------------------------------------------------------- runtime state
cudaStream_t stream();        // stream every kernel of this library is launched on
int          ensure_device(); // lazily binds the device; AB_ECUDA if there is no usable GPU
int          sm_count();
void         count_launch();
double       opt(const char* key, double dflt);
void         timer_add(int which, double seconds);
// Launch on the library stream; with the option "pdl" = 1 the launch carries the programmatic-stream-serialization
// attribute, so the kernel's prologue overlaps the tail of the previous one (the kernel must call abd::pdl_wait()).
// OFF by default — measured (profiles/slab_floor_r02p.log): 564.6 vs 579.0 CG it/s at 512^3, 4114 vs 4187 on the N=8 slab;
// the per-kernel times are unchanged, the early-resident successor CTAs cost the running kernel more than the hidden
// launch latency returns.
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream();
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = opt("pdl", 0.0) != 0.0 ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(static_cast<Args&&>(args))...);
}

// ---------------------------------------------------------------- device memory
int dev_alloc(void** p, size_t bytes);
int dev_free(void* p);

template <class T>
struct DevBuf {
    T*     p = nullptr;
    size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf&)            = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    int alloc(size_t count) {
        release();
        n = count;
        if (count == 0) count = 1;
        return dev_alloc((void**)&p, count * sizeof(T));
    }
    void release() {
        if (p) dev_free(p);
        p = nullptr;
        n = 0;
    }
    T* take() {
        T* q = p;
        p    = nullptr;
        n    = 0;
        return q;
    }
};

bool is_device_ptr(const void* p);

// Input argument that may live on host or device: dev() is always a device pointer.
template <class T>
struct In {
    const T* d = nullptr;
    DevBuf<T> tmp;
    int bind(const T* p, size_t count) {
        if (count == 0 || p == nullptr) { d = p; return AB_OK; }
        if (is_device_ptr(p)) { d = p; return AB_OK; }
        AB_TRY(tmp.alloc(count));
        AB_CUDA(cudaMemcpyAsync(tmp.p, p, count * sizeof(T), cudaMemcpyHostToDevice, stream()));
        d = tmp.p;
        return AB_OK;
    }
};

// Output argument that may live on host or device; commit() copies back when host.
template <class T>
struct Out {
    T*  d    = nullptr;
    T*  host = nullptr;
    size_t n = 0;
    DevBuf<T> tmp;
    int bind(T* p, size_t count) {
        n = count;
        if (count == 0 || p == nullptr) { d = p; return AB_OK; }
        if (is_device_ptr(p)) { d = p; return AB_OK; }
        AB_TRY(tmp.alloc(count));
        d    = tmp.p;
        host = p;
        return AB_OK;
    }
    // asynchronous copy-back followed by a stream sync (host memory must be valid on return)
    int commit() {
        if (host && n) {
            AB_CUDA(cudaMemcpyAsync(host, d, n * sizeof(T), cudaMemcpyDeviceToHost, stream()));
            AB_CUDA(cudaStreamSynchronize(stream()));
        }
        return AB_OK;
    }
};

// ---------------------------------------------------------------- CSR handle
constexpr int CSR_PAD = 32;   // entries of padding after colind/values: aligned bulk copies over-read
struct Csr {
    int64_t m = 0, n = 0, nnz = 0, T = 0;
    // CSR proper. colind/values are padded by CSR_PAD entries so aligned bulk copies may over-read.
    int32_t* rowptr = nullptr;   // [m+1]
    int32_t* colind = nullptr;   // [nnz+CSR_PAD]
    double*  values = nullptr;   // [nnz+CSR_PAD]
    // offset-dictionary form of colind for the SpMV tile kernel (lazy, spmv.cu): off8[e] indexes
    // offtab[256] = sorted distinct (col - row); noff: -1 not attempted, 0 unavailable, else count
    uint8_t* off8   = nullptr;   // [nnz+64]
    int32_t* offtab = nullptr;   // [256]
    int      noff   = -1;
    // pattern-code form for the iterative solvers (lazy, spmv.cu): when the value array a solver
    // multiplies with (the scaled copy svalues) has so few distinct (offset, value) pairs that one
    // byte names the pair, code8[e] indexes codetab[256] = {value, offset}: 1 instead of 9 bytes per
    // stored entry.  ncode: -1 not attempted, 0 unavailable, else number of pairs.
    uint8_t* code8   = nullptr;    // [nnz+64]
    void*    codetab = nullptr;    // [256] x {double value; int32 offset; int32 pad}
    int      ncode   = -1;
    uint64_t code_version = 0;     // values_version the codes were built from
    const double* code_src = nullptr;   // the value array they encode
    // row-pattern form on top of the pattern codes (lazy, spmv.cu): when the rows' code sequences
    // come from at most 256 distinct patterns (a 3-D 7-point stencil has 27), rowcode[r] names the
    // pattern and pat_ent / pat_info hold the patterns: 1 byte per ROW instead of per entry, no rowptr.
    uint8_t*  rowcode  = nullptr;  // [m + 1024] (zero padded: 16-byte bulk copies may over-read)
    void*     pat_ent  = nullptr;  // [npat_ent] x {double value; int32 offset; int32 pad}, patterns back to back
    uint32_t* pat_info = nullptr;  // [256] start | len << 16
    int       npat = -1, npat_ent = 0;   // npat: -1 not attempted, 0 unavailable
    // x-window form of the row patterns (lazy, spmv.cu: spmv_rowwin_kernel): a 1024-row tile needs x only
    // inside a few index intervals (one per cluster of stencil offsets); the TMA unit stages those intervals
    // in shared memory and the pattern entries address the stage directly.  rw_ok: -1 not attempted, 0 n/a.
    void*     rw_plan  = nullptr;  // host RowWinPlan (spmv.cu)
    void*     rw_ent   = nullptr;  // [npat_ent] x {double value; int32 stage index rel. to the local row; int32 pad}
    uint8_t*  rw_tmask = nullptr;  // [ceil(m/1024)] bit i: interval i is referenced by a row of the tile
    int       rw_ok = -1;
    // marching form (lazy, spmv.cu: spmv_rowmarch_kernel): chains of tiles M rows apart share their windows
    void*     rm_plan  = nullptr;  // host RowMarchPlan
    void*     rm_ent   = nullptr;  // [npat_ent] x {double value; int32 offset inside the plane buffer; int32 plane}
    int32_t*  rm_items = nullptr;  // [2 * rm_nitems]: first tile of each work item, then its length
    int       rm_nitems = 0;
    int64_t   rm_nrest = 0;        // tiles no chain covers (they reference other intervals, or are excluded)
    int32_t*  rm_rest = nullptr;   // [rm_nrest] those tiles, ascending
    int       rm_ok = 0;
    void*     zm_plan  = nullptr;  // host copy of the z-register marching plan (spmv.cu ZMarchPlan)
    void*     zm_ent   = nullptr;  // [npat_ent] x {double value; int32 offset inside the window; int32 plane}
    int32_t*  zm_items = nullptr;  // [3 * zm_nitems]: first row of each work item's first chain position, its length, rows per position
    int       zm_nitems = 0;
    int32_t*  zm_cta_first = nullptr; // [zm_ncta + 1] balanced partition: CTA b runs items [zm_cta_first[b], zm_cta_first[b+1]) (null: round-robin)
    int       zm_ncta = 0;
    int64_t   zm_nrest = 0;        // 1024-row tiles no item covers
    int32_t*  zm_rest = nullptr;   // [zm_nrest] those tiles, ascending
    int       zm_ok = 0;
    uint8_t*  zm_rowmask = nullptr; // [m + pad] zm_mask[rowcode[r]]: the byte stream the kernel stages when the dominant pattern is in its parameters
    uint8_t*  zm_mask = nullptr;   // [256] per pattern: which entries of the dominant pattern it keeps (0x80: walk the table)
    uint8_t*  rw_mask = nullptr;     // [256] per pattern: which entries of the window kernel's dominant pattern it keeps (0x80: walk)
    uint8_t*  rw_exclude = nullptr;  // [ceil(m/1024)] optional (dist.cu): tiles that must not join a chain (they read halo columns)
    // triplet bookkeeping (null when the handle was created from CSR: identity map)
    uint32_t* perm     = nullptr; // [T]  sorted position -> input triplet
    uint32_t* segstart = nullptr; // [nnz+1] first sorted position of each stored entry
    int32_t*  slot     = nullptr; // [T]  input triplet -> stored entry
    int32_t*  entry_row = nullptr; // [nnz] row of each stored entry (lazy)
    // SpMV tiling facts (computed at finalize)
    int32_t max_tile_nnz[3] = {0, 0, 0};   // max nnz over 256- / 128- / 64-row tiles
    int32_t max_row_nnz  = 0;
    // Jacobi: 1/diag (1 where the diagonal is absent or zero), lazy
    double* dinv = nullptr;
    bool    dinv_valid = false;
    // symmetrically scaled copy for CG (lazy): svalues = S A S, sinv = diag(A)^-1/2
    double*  svalues = nullptr;    // [nnz+8]
    double*  sinv    = nullptr;    // [m (+ halo)]
    uint64_t values_version = 1;   // bumped whenever `values` change
    uint64_t scaled_version = 0;   // version svalues/sinv were built from (0 = never)
    bool     scaled_ok = false;    // false: a diagonal entry is missing / non-positive -> classic path
    // numerical symmetry A == A' (bitwise), checked lazily per values_version (solver.cu): lets the
    // legacy direct-solver strings try CG first and lets the adjoint solve skip the transpose
    uint64_t sym_version = 0;
    bool     is_sym = false;
    // cached transpose (lazy): handle id of A', and map tperm (A' entry -> A entry)
    int64_t   th    = 0;
    uint32_t* tperm = nullptr;
    // solver workspace (lazy): 8 vectors of length max(m,n) + scalars
    double* work = nullptr;
    size_t  work_len = 0;
    void*   scal = nullptr;       // device SolveScal block (solver.cu)
    double* partials = nullptr;   // reduction partials
    uint32_t* ticket = nullptr;   // last-block tickets
    void* graph_cache = nullptr;
    // long-row work list for SpMM / SDDMM (lazy): rows longer than LONG_ROW are split into
    // CHUNK-entry pieces so one power-law row cannot serialise a whole warp
    int32_t* long_rows = nullptr;         // [long_nrows]
    int32_t* long_row_chunk0 = nullptr;   // [long_nrows+1] first chunk of each long row
    int32_t* long_chunk_row = nullptr;    // [long_nchunks]
    int32_t* long_chunk_start = nullptr;  // [long_nchunks] first entry of the chunk
    int64_t  long_nrows = -1, long_nchunks = 0;   // -1: not built yet
    ~Csr();
};

Csr*    csr_get(int64_t h);       // nullptr + error set when unknown
int64_t csr_put(Csr* c);
int     csr_erase(int64_t h);
int     csr_finalize(Csr* c);     // computes tiling facts after rowptr/colind/values are in place
int     csr_ensure_dinv(Csr* c);
int     csr_ensure_entry_row(Csr* c);
int     csr_ensure_slot(Csr* c);  // lazy triplet -> stored-entry map (null afterwards = identity)
int     csr_ensure_transpose(Csr* c, Csr** out);
int     csr_ensure_work(Csr* c, size_t nvec);
std::mutex& big_lock();

// ---------------------------------------------------------------- primitives (assemble.cu)
// exclusive prefix sum of uint32 (in -> out, may alias); total written to *d_total (device) if non-null
int scan_exclusive_u32(const uint32_t* in, uint32_t* out, size_t n, uint32_t* d_total);
// stable LSD radix sort of 64-bit keys by their bits [bit0, bit0 + nbits), carrying a 32-bit payload that the first pass
// synthesises from the position (the payload on return is the input position of each element)
int radix_sort_u64(uint64_t* keys, uint32_t* vals, uint64_t* keys_tmp, uint32_t* vals_tmp,
                   size_t n, int nbits, uint64_t** keys_out, uint32_t** vals_out, int bit0 = 0);

// ---------------------------------------------------------------- spmv.cu
// y = A x.  When dotvec != nullptr the launch also produces dot_result[0] = <y, dotvec> and
// dot_result[1] = <y, y> (deterministic two-stage reduction through partials/ticket).
// done_flag (device int, may be null): the kernel is a no-op when *done_flag != 0.
constexpr int MAX_REDUCE_GRID = 8192;   // partials buffers hold NV * MAX_REDUCE_GRID doubles
constexpr int TICKET_GROUP = 32;        // CTAs per first-level ticket word (abd::grid_reduce)
constexpr int TICKET_WORDS = 1 + MAX_REDUCE_GRID / TICKET_GROUP;   // every ticket buffer holds this many zeroed uint32
int launch_spmv(const Csr* c, const double* x, double* y, const double* dotvec, double* dot_result,
                double* partials, uint32_t* ticket, const int* done_flag);
// restricted to a list of tiles (distributed SpMV: interior first, boundary after the halo);
// accumulate != 0 adds the fused dots onto dot_result.  spmv_tile_rows: rows per tile, 0 = n/a.
int launch_spmv_part(const Csr* c, const double* x, double* y, const double* dotvec, double* dot_result,
                     double* partials, uint32_t* ticket, const int* done_flag, const int32_t* tile_list,
                     int64_t nlist, int accumulate);
int spmv_tile_rows(const Csr* c);
// Distributed SpMV in ONE launch: the tile list holds the interior tiles first; before a CTA touches
// list position >= n_before it waits (acquire) until every peer with recv_cnt > 0 has published `seq`
// in flags[peer].  Only the row-pattern kernel implements it (spmv_supports_halo_wait()).
constexpr int P2P_MAXW = 16;   // ranks on one NVSwitch box
// Lives in cudaMalloc memory on each rank and is IPC-mapped by every peer (dist.cu).
struct P2PBlock {
    double   red[2][P2P_MAXW][2];     // all-reduce slots [parity][source rank][2 values]
    uint32_t red_flag[2][P2P_MAXW];   // sequence number whose values sit in red[parity][source]
    uint32_t halo_flag[P2P_MAXW];     // halo_flag[source] = sequence number of its last finished push
    uint32_t err;                     // a spin loop timed out
    uint32_t ticket;                  // last-CTA ticket of halo_push_kernel
    // box-wide sums re-published LOCALLY by the one CTA that polls the remotely written red_flag words, so that
    // the other CTAs of a kernel wait on a device-scope flag instead of all hammering the NVLink-written lines
    double   sum[2][2];               // [parity][2 values]
    uint32_t sum_flag[2];             // sequence number whose sums sit in sum[parity]
    // flag-in-payload slots of p2p_allreduce2_kernel: four 8-byte words per (parity, source rank), each = (sequence
    // number << 32) | one 32-bit half of the two doubles — an aligned 8-byte store arrives whole, so a word whose upper
    // half shows the awaited sequence number carries its data and no fence / separate flag round trip is needed
    unsigned long long ll[2][P2P_MAXW][4];
};
struct HaloWait {
    const uint32_t* flags;     // [world] sequence numbers published by the peers (this rank's memory)
    const int*      recv_cnt;  // [world]
    int             world;
    uint32_t        seq;
    unsigned        n_before;  // list positions below this need no halo
    uint32_t*       err;       // set when a wait times out
    // optional: the kernel's last CTA pushes the two fused dot products straight into every peer's
    // all-reduce slots (sequence red_seq) instead of leaving them for a separate all-reduce launch
    P2PBlock* const* peer_blk; // [world] (nullptr: no push)
    int              me;
    uint32_t         red_seq;
    // exchange != 0: the last CTA does the WHOLE all-reduce (tagged words to every peer, poll, rank-order sum) and leaves
    // the box-wide sums in dot_result — no all-reduce launch follows the SpMV
    P2PBlock*        my_blk;
    int              exchange;
};
// the same for a vector kernel with a fused inner product (cgs_update_kernel): blk == nullptr: no exchange
struct ArExch {
    P2PBlock*        blk = nullptr;
    P2PBlock* const* peer_blk = nullptr;
    int              world = 1, me = 0;
    uint32_t         seq = 0;
};
bool spmv_supports_halo_wait(const Csr* c, const double* vals);
int  launch_spmv_halo(const Csr* c, const double* vals, const double* x, double* y, const double* dotvec,
                      double* dot_result, double* partials, uint32_t* ticket, const int* done_flag,
                      const int32_t* tile_list, int64_t nlist, const HaloWait& hw, int list_rows);
// Distributed SpMV with the marching kernel: chains (tiles that read no halo column) first, then — after the
// in-kernel acquire of the peers' halo flags — the remaining tiles through the x-window kernel, which adds its share
// of the fused dots and, when hw.peer_blk is set, pushes the totals to the peers.  false: not available for (c, vals, x).
bool spmv_march_halo_ready(const Csr* c, const double* vals, const double* x);
int  launch_spmv_march_halo(const Csr* c, const double* vals, const double* x, double* y, const double* dotvec, double* dot_result,
                            double* partials, uint32_t* ticket, const int* done_flag, const HaloWait& hw);
// rows per tile that launch_spmv_halo expects in its tile list for this matrix / operand (0: unsupported)
int  spmv_halo_tile_rows(const Csr* c, const double* vals, const double* x);
constexpr int SPMV_WINDOW_TILE_ROWS = 1024;   // tile size of the x-window SpMV kernel (spmv.cu RW_R)
// same with an explicit value array (c->values or the scaled copy c->svalues)
int launch_spmv_vals(const Csr* c, const double* vals, const double* x, double* y, const double* dotvec,
                     double* dot_result, double* partials, uint32_t* ticket, const int* done_flag,
                     const int32_t* tile_list, int64_t nlist, int accumulate);
// builds / refreshes c->svalues and c->sinv; c->scaled_ok tells whether the scaled CG may be used
int csr_ensure_scaled(Csr* c);
int launch_spmm(const Csr* c, const double* X, int64_t k, double* Y);

}  // namespace ab

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
namespace abd {

// ---- programmatic dependent launch (griddepcontrol): a kernel launched through ab::launch_pdl may become resident
// while its predecessor in the stream is still running; pdl_wait() returns once that predecessor has completed and
// its writes are visible, pdl_launch_dependents() lets the successor's CTAs in as soon as resources allow.  Both are
// no-ops for a kernel launched the ordinary way.  What a kernel does BEFORE pdl_wait() must not depend on — or
// disturb — anything an earlier kernel touches (table loads of static data, barrier set-up).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum; result valid in thread 0. `sh` must hold >= 32 doubles.
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* sh) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double r = 0.0;
    if (w == 0) {
        r = (lane < THREADS / 32) ? sh[lane] : 0.0;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

// Deterministic grid reduction: each block stores its partial, the last block to arrive
// (ticket counter) sums all partials in a fixed order and calls fin(total).
// Must be called by all threads of the block; `mine` valid in thread 0.
template <int THREADS, int NV, class Fin>
__device__ __forceinline__ void grid_reduce(const double (&mine)[NV], double* partials,
                                            uint32_t* ticket, double* sh, Fin fin) {
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) partials[(size_t)k * gridDim.x + blockIdx.x] = mine[k];
        __threadfence();
        // two-level ticket: the CTAs of a kernel finish together, and several hundred atomics on ONE word serialise in the
        // L2 (tens of cycles each); groups of 32 CTAs count on their own word, the last of each group on word 0
        const unsigned g = blockIdx.x / ab::TICKET_GROUP, ng = (gridDim.x + ab::TICKET_GROUP - 1) / ab::TICKET_GROUP;
        const unsigned gsize = min((unsigned)ab::TICKET_GROUP, gridDim.x - g * ab::TICKET_GROUP);
        bool last = false;
        if (atomicAdd(ticket + 1 + g, 1u) == gsize - 1) {
            ticket[1 + g] = 0;                    // re-armed for the next launch (nobody else touches it in this one)
            __threadfence();
            last = atomicAdd(ticket, 1u) == ng - 1;
        }
        s_last = last;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        double tot[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double a = 0.0;
            const volatile double* pk = partials + (size_t)k * gridDim.x;
            for (unsigned i = threadIdx.x; i < gridDim.x; i += THREADS) a += pk[i];
            tot[k] = block_sum<THREADS>(a, sh);
        }
        if (threadIdx.x == 0) {
            *ticket = 0;   // re-arm for the next launch
            fin(tot);
        }
    }
}

// ---- mbarrier + 1-D bulk async copy (TMA unit, SASS: UBLKCP) -------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; dst, src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ---- peer-memory primitives (system-scope release/acquire over NVLink-mapped memory) -------------
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// spins until *p has reached seq.  Every collective entry point opens with a stream-ordered NCCL rendezvous
// (dist.cu: dist_entry_barrier), so these spins only ever cover the skew between ranks that are already inside
// the same call; the bound (2^36 SM clocks, about half a minute) exists so that a rank that DIED mid-solve turns
// into an error (AB_ENCCL through *err) on its peers instead of a hung GPU.
constexpr int AB_WAIT_CLOCKS_LOG2 = 36;
__device__ __forceinline__ void wait_seq(const uint32_t* p, uint32_t seq, uint32_t* err) {
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(p) - seq) < 0) {
        if (clock64() - t0 > (1ll << AB_WAIT_CLOCKS_LOG2)) { atomicExch(err, 1u); break; }
        __nanosleep(20);
    }
}
// one thread: deposit two partial sums in every rank's slot (own included) and publish `seq`
__device__ __forceinline__ void p2p_push2(ab::P2PBlock* const* peer_blk, int world, int me, uint32_t seq, double v0, double v1) {
    const int par = seq & 1u;
    for (int p = 0; p < world; ++p) {
        ab::P2PBlock* dst = peer_blk[p];
        dst->red[par][me][0] = v0;
        dst->red[par][me][1] = v1;
    }
    __threadfence_system();
    for (int p = 0; p < world; ++p) st_release_sys(&peer_blk[p]->red_flag[par][me], seq);
}
// ONE thread: box-wide sums of (v0, v1) by the flag-in-payload exchange (P2PBlock::ll) — four tagged 8-byte words to every
// rank's slot for `me` (own included), then the eight slots of this rank polled and added in rank order: same bits on
// every rank.  Called from the last CTA of the kernel that produced the partial sums, so no all-reduce launch follows.
__device__ __forceinline__ void ll_allreduce2(ab::P2PBlock* blk, ab::P2PBlock* const* peer_blk, int world, int me, uint32_t seq,
                                              double& v0, double& v1) {
    const int par = seq & 1u;
    const unsigned long long b0 = (unsigned long long)__double_as_longlong(v0), b1 = (unsigned long long)__double_as_longlong(v1);
    const unsigned long long tag = (unsigned long long)seq << 32;
    for (int p = 0; p < world; ++p) {
        unsigned long long* dst = &peer_blk[p]->ll[par][me][0];
        asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst + 0), "l"(tag | (b0 & 0xffffffffull)) : "memory");
        asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst + 1), "l"(tag | (b0 >> 32)) : "memory");
        asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst + 2), "l"(tag | (b1 & 0xffffffffull)) : "memory");
        asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst + 3), "l"(tag | (b1 >> 32)) : "memory");
    }
    double s0 = 0.0, s1 = 0.0;
    const long long t0 = clock64();
    for (int p = 0; p < world; ++p) {
        const unsigned long long* src = &blk->ll[par][p][0];
        unsigned long long w[4];
        for (;;) {
            bool ok = true;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w[i]) : "l"(src + i) : "memory");
                ok = ok && (uint32_t)(w[i] >> 32) == seq;
            }
            if (ok) break;
            if (clock64() - t0 > (1ll << AB_WAIT_CLOCKS_LOG2)) { atomicExch(&blk->err, 1u); break; }
        }
        s0 += __longlong_as_double((long long)((w[1] << 32) | (w[0] & 0xffffffffull)));
        s1 += __longlong_as_double((long long)((w[3] << 32) | (w[2] & 0xffffffffull)));
    }
    v0 = s0;
    v1 = s1;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// whole CTA, every CTA of the grid (all of them must be co-resident): the box-wide sums of contribution `seq`.
// CTA 0 alone waits for the eight remotely written flags (system scope), adds the slots in rank order (same bits on
// every rank) and re-publishes the two sums in local memory; every other CTA waits for that local flag (device
// scope, one L2 line).  Measured motivation: with all 1184 CTAs polling the NVLink-written flag words the folded-in
// all-reduce LOST to separate all-reduce launches at 8 ranks (2855 vs 2994 it/s).  sh2 = 2 doubles of shared memory.
__device__ __forceinline__ void p2p_wait_sum2(ab::P2PBlock* blk, int world, uint32_t seq, double* sh2, double& s0, double& s1) {
    const int par = seq & 1u;
    if (blockIdx.x == 0) {
        if ((int)threadIdx.x < world) wait_seq(&blk->red_flag[par][threadIdx.x], seq, &blk->err);
        __syncthreads();
        if (threadIdx.x == 0) {
            double a0 = 0.0, a1 = 0.0;
            const volatile double* r = &blk->red[par][0][0];
            for (int p = 0; p < world; ++p) { a0 += r[2 * p]; a1 += r[2 * p + 1]; }
            blk->sum[par][0] = a0; blk->sum[par][1] = a1;
            __threadfence();
            st_release_gpu(&blk->sum_flag[par], seq);
            sh2[0] = a0; sh2[1] = a1;
        }
    } else if (threadIdx.x == 0) {
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_gpu(&blk->sum_flag[par]) - seq) < 0) {
            if (clock64() - t0 > (1ll << AB_WAIT_CLOCKS_LOG2)) { atomicExch(&blk->err, 1u); break; }
            __nanosleep(40);
        }
        const volatile double* sm = &blk->sum[par][0];
        sh2[0] = sm[0]; sh2[1] = sm[1];
    }
    __syncthreads();
    s0 = sh2[0]; s1 = sh2[1];
}

__device__ __forceinline__ double2 ldg_stream_d2(const double2* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

}  // namespace abd
#endif
