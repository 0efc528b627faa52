// dist.cu — row-block distributed CSR, halo exchange and Jacobi-PCG across the GPUs of one box.
//
// Replaces the reference's mpi_SparseTensor path (src/mpi.jl:420-508; MPITensorSolve_forward,
// deps/Plugin/MPITensor/MPITensor.h:103-152 = Hypre IJ assemble + BoomerAMG/GMRES over MPI):
// one PROCESS per GPU; rank r owns the contiguous global rows [ilower, iupper] (Hypre's IJ
// layout, inclusive) and holds them as a local CSR with GLOBAL column ids.
//
// Setup (once per pattern): the off-rank columns a rank touches are found on the device
// (flag -> compact -> radix sort -> unique), the local CSR is re-indexed to
// [0,nloc) = owned columns, [nloc, nloc+next) = halo columns, and the owners learn which of
// their rows to ship (index lists swapped with ncclSend/ncclRecv).
// SpMV: pack owned boundary values -> grouped ncclSend/ncclRecv straight into the halo tail of
// the extended vector (NVLink P2P under NCCL) -> the same TMA-staged SpMV kernel as one GPU.
// CG: the two inner products per iteration are 2-element fp64 ncclAllReduce calls on the same
// stream; the scalar recurrences stay on the device (cg_fin_kernel), no host round trip.
//
// NCCL is dlopen()ed (libnccl.so.2 — the copy PyTorch already loaded when the host is Python),
// so single-GPU users never need it.
#include "common.cuh"
#include "cg_kernels.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <map>
#include <string>
#include <algorithm>

namespace ab {

struct Nccl {
    void* so = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static Nccl g_nccl;
static ncclComm_t g_comm = nullptr;
static int g_rank = 0, g_world = 1;

static int nccl_load() {
    if (g_nccl.so) return AB_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* so = nullptr;
    for (const char* nm : names) { so = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (so) break; }
    if (!so) return fail(AB_ENCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define AB_SYM(field, name)                                                   \
    *(void**)(&g_nccl.field) = dlsym(so, name);                               \
    if (!g_nccl.field) return fail(AB_ENCCL, "libnccl lacks symbol %s", name)
    AB_SYM(GetUniqueId, "ncclGetUniqueId");
    AB_SYM(CommInitRank, "ncclCommInitRank");
    AB_SYM(CommDestroy, "ncclCommDestroy");
    AB_SYM(AllReduce, "ncclAllReduce");
    AB_SYM(AllGather, "ncclAllGather");
    AB_SYM(Send, "ncclSend");
    AB_SYM(Recv, "ncclRecv");
    AB_SYM(GroupStart, "ncclGroupStart");
    AB_SYM(GroupEnd, "ncclGroupEnd");
    AB_SYM(GetErrorString, "ncclGetErrorString");
#undef AB_SYM
    g_nccl.so = so;
    return AB_OK;
}

#define AB_NCCL(call)                                                                          \
    do {                                                                                       \
        ncclResult_t r__ = (call);                                                             \
        if (r__ != ncclSuccess)                                                                \
            return ab::fail(AB_ENCCL, "%s:%d %s -> %s", __FILE__, __LINE__, #call,             \
                            g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "nccl error"); \
    } while (0)

struct DistCsr {
    Csr* loc = nullptr;            // re-indexed local block: m = nloc, n = nloc + next
    int64_t ilower = 0, iupper = -1, N = 0;
    int64_t nloc = 0, next = 0;
    std::vector<int> send_cnt, send_off, recv_cnt, recv_off;   // per peer, in vector entries
    int64_t nsend = 0;
    int32_t* send_idx = nullptr;   // [nsend] local row ids to ship, grouped by destination
    double*  sendbuf  = nullptr;   // [nsend]
    double*  xext     = nullptr;   // [nloc + next (+pad)]  SpMV operand / CG direction vector
    ~DistCsr() { delete loc; dev_free(send_idx); dev_free(sendbuf); dev_free(xext); }
};
static std::map<int64_t, DistCsr*> g_dist;
static int64_t g_next_dh = 1;
static std::mutex g_dist_mu;

static DistCsr* dist_get(int64_t dh) {
    std::lock_guard<std::mutex> lk(g_dist_mu);
    auto it = g_dist.find(dh);
    if (it == g_dist.end()) { set_error("unknown distributed handle %lld", (long long)dh); return nullptr; }
    return it->second;
}

// ---------------------------------------------------------------- setup kernels
__global__ void ext_flag_kernel(const int32_t* __restrict__ colind, size_t nnz, int64_t lo, int64_t hi,
                                uint32_t* __restrict__ flags) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < nnz) { int64_t cidx = colind[e]; flags[e] = (cidx < lo || cidx > hi) ? 1u : 0u; }
}
__global__ void ext_compact_kernel(const int32_t* __restrict__ colind, size_t nnz, int64_t lo, int64_t hi,
                                   const uint32_t* __restrict__ excl, uint64_t* __restrict__ keys) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < nnz) { int64_t cidx = colind[e]; if (cidx < lo || cidx > hi) keys[excl[e]] = (uint64_t)cidx; }
}
__global__ void uniq_flag_kernel(const uint64_t* __restrict__ keys, size_t n, uint32_t* __restrict__ flags) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) flags[k] = (k == 0 || keys[k] != keys[k - 1]) ? 1u : 0u;
}
__global__ void uniq_emit_kernel(const uint64_t* __restrict__ keys, size_t n, const uint32_t* __restrict__ excl,
                                 int64_t* __restrict__ ext) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n && (k == 0 || keys[k] != keys[k - 1])) ext[excl[k]] = (int64_t)keys[k];
}
__global__ void remap_cols_kernel(const int32_t* __restrict__ colin, size_t nnz, int64_t lo, int64_t hi, int64_t nloc,
                                  const int64_t* __restrict__ ext, int64_t next, int32_t* __restrict__ colout) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    int64_t cidx = colin[e];
    if (cidx >= lo && cidx <= hi) { colout[e] = (int32_t)(cidx - lo); return; }
    int64_t a = 0, b = next;   // lower_bound
    while (a < b) { int64_t mid = (a + b) >> 1; if (ext[mid] < cidx) a = mid + 1; else b = mid; }
    colout[e] = (int32_t)(nloc + a);
}
__global__ void to_local_idx_kernel(const int64_t* __restrict__ gidx, size_t n, int64_t lo, int64_t hi,
                                    int32_t* __restrict__ lidx, int* __restrict__ err) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int64_t g = gidx[k];
    if (g < lo || g > hi) { atomicExch(err, 1); g = lo; }
    lidx[k] = (int32_t)(g - lo);
}
__global__ void pack_kernel(const int32_t* __restrict__ idx, const double* __restrict__ x, size_t n, double* __restrict__ buf) {
    size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) buf[k] = x[idx[k]];
}

static int halo_exchange(DistCsr* d) {
    if (g_world == 1 || (d->nsend == 0 && d->next == 0)) return AB_OK;
    if (d->nsend) {
        pack_kernel<<<(unsigned)((d->nsend + 255) / 256), 256, 0, stream()>>>(d->send_idx, d->xext, (size_t)d->nsend, d->sendbuf);
        AB_LAUNCH_CHECK();
    }
    AB_NCCL(g_nccl.GroupStart());
    for (int p = 0; p < g_world; ++p) {
        if (p == g_rank) continue;
        if (d->send_cnt[p]) AB_NCCL(g_nccl.Send(d->sendbuf + d->send_off[p], (size_t)d->send_cnt[p], ncclDouble, p, g_comm, stream()));
        if (d->recv_cnt[p]) AB_NCCL(g_nccl.Recv(d->xext + d->nloc + d->recv_off[p], (size_t)d->recv_cnt[p], ncclDouble, p, g_comm, stream()));
    }
    AB_NCCL(g_nccl.GroupEnd());
    return AB_OK;
}

static int allreduce2(double* p) {
    if (g_world == 1) return AB_OK;
    AB_NCCL(g_nccl.AllReduce(p, p, 2, ncclDouble, ncclSum, g_comm, stream()));
    return AB_OK;
}

// CG over the row-block partition.  b, x: local device vectors of length nloc.  tol < 0: fixed count.
static int dist_run_cg(DistCsr* d, const double* b, double* x, double tol, int maxit, SolveScal* fin) {
    Csr* c = d->loc;
    const size_t n = (size_t)d->nloc;
    AB_TRY(csr_ensure_dinv(c));
    AB_TRY(csr_ensure_work(c, 4));
    double *r = work_vec(c, 0), *Ap = work_vec(c, 1), *p = d->xext;
    SolveScal* s = (SolveScal*)c->scal;
    const unsigned g = ew_grid(n);
    const int raw = g_world > 1 ? 1 : 0;
    cg_init_kernel<<<g, EW_THREADS, 0, stream()>>>(b, c->dinv, x, r, p, n, tol, s, c->partials, c->ticket, raw);
    AB_LAUNCH_CHECK();
    if (raw) {
        AB_TRY(allreduce2(s->red));
        cg_fin_kernel<<<1, 1, 0, stream()>>>(s, 0, 0, tol);
        AB_LAUNCH_CHECK();
    }
    int check = (int)opt("check_every", 0);
    if (check <= 0) check = 16;
    int it = 0;
    SolveScal h;
    while (it < maxit) {
        int chunk = maxit - it < check ? maxit - it : check;
        for (int k = 0; k < chunk; ++k, ++it) {
            const int parity = it & 1;
            AB_TRY(halo_exchange(d));
            AB_TRY(launch_spmv(c, p, Ap, p, s->dots, c->partials, c->ticket, &s->done));
            if (raw) AB_TRY(allreduce2(s->dots));
            cg_update_kernel<<<g, EW_THREADS, 0, stream()>>>(p, Ap, c->dinv, x, r, n, parity, s, c->partials, c->ticket, raw);
            AB_LAUNCH_CHECK();
            if (raw) {
                AB_TRY(allreduce2(s->red));
                cg_fin_kernel<<<1, 1, 0, stream()>>>(s, 1, parity, tol);
                AB_LAUNCH_CHECK();
            }
            cg_direction_kernel<<<g, EW_THREADS, 0, stream()>>>(r, c->dinv, p, n, parity, s);
            AB_LAUNCH_CHECK();
        }
        if (tol < 0 && it < maxit) continue;
        AB_TRY(read_scal(c, &h));
        if (h.done) break;   // identical on every rank: the flags derive from all-reduced values
    }
    AB_TRY(read_scal(c, fin));
    return AB_OK;
}

}  // namespace ab

using namespace ab;

extern "C" {

int ab_dist_unique_id(void* id128) {
    if (!id128) return fail(AB_EINVAL, "null id buffer");
    AB_TRY(nccl_load());
    ncclUniqueId id;
    AB_NCCL(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    memcpy(id128, &id, 128);
    return AB_OK;
}

int ab_dist_init(int rank, int world, const void* id128) {
    AB_TRY(ensure_device());
    if (world < 1 || rank < 0 || rank >= world) return fail(AB_EINVAL, "bad rank/world");
    g_rank = rank; g_world = world;
    if (world == 1) return AB_OK;
    if (!id128) return fail(AB_EINVAL, "null id buffer");
    AB_TRY(nccl_load());
    if (g_comm) { g_nccl.CommDestroy(g_comm); g_comm = nullptr; }
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    AB_NCCL(g_nccl.CommInitRank(&g_comm, world, id, rank));
    return AB_OK;
}

int ab_dist_finalize(void) {
    std::vector<DistCsr*> all;
    {
        std::lock_guard<std::mutex> lk(g_dist_mu);
        for (auto& kv : g_dist) all.push_back(kv.second);
        g_dist.clear();
    }
    for (auto* d : all) delete d;
    if (g_comm && g_nccl.CommDestroy) { g_nccl.CommDestroy(g_comm); g_comm = nullptr; }
    g_rank = 0; g_world = 1;
    return AB_OK;
}

int ab_dist_csr_create(int64_t h_local, int64_t ilower, int64_t iupper, int64_t N, int64_t* out_dh) {
    AB_TRY(ensure_device());
    Csr* src = csr_get(h_local);
    if (!src) return AB_EHANDLE;
    if (!out_dh) return fail(AB_EINVAL, "out_dh is null");
    const int64_t nloc = iupper - ilower + 1;
    if (nloc != src->m || src->n != N || ilower < 0 || iupper >= N)
        return fail(AB_EINVAL, "row range [%lld,%lld] does not match the local block (%lld x %lld), N=%lld",
                    (long long)ilower, (long long)iupper, (long long)src->m, (long long)src->n, (long long)N);
    if (g_world > 1 && !g_comm) return fail(AB_ENCCL, "ab_dist_init has not been called");
    std::lock_guard<std::mutex> lk(big_lock());
    DistCsr* d = new DistCsr();
    d->ilower = ilower; d->iupper = iupper; d->N = N; d->nloc = nloc;
    const size_t nnz = (size_t)src->nnz;
    auto body = [&]() -> int {
        // ---- 1. every rank's row range
        std::vector<int64_t> ranges(2 * (size_t)g_world);
        ranges[2 * g_rank] = ilower; ranges[2 * g_rank + 1] = iupper;
        if (g_world > 1) {
            DevBuf<int64_t> dr;
            AB_TRY(dr.alloc(2 * (size_t)g_world));
            AB_CUDA(cudaMemcpyAsync(dr.p + 2 * g_rank, &ranges[2 * g_rank], 2 * sizeof(int64_t), cudaMemcpyHostToDevice, stream()));
            AB_NCCL(g_nccl.AllGather(dr.p + 2 * g_rank, dr.p, 2, ncclInt64, g_comm, stream()));
            AB_CUDA(cudaMemcpyAsync(ranges.data(), dr.p, ranges.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, stream()));
            AB_CUDA(cudaStreamSynchronize(stream()));
        }
        // ---- 2. sorted unique off-rank columns
        DevBuf<int64_t> ext;
        uint32_t next = 0;
        if (nnz) {
            DevBuf<uint32_t> flags, total;
            AB_TRY(flags.alloc(nnz)); AB_TRY(total.alloc(1));
            unsigned nb = (unsigned)((nnz + 255) / 256);
            ext_flag_kernel<<<nb, 256, 0, stream()>>>(src->colind, nnz, ilower, iupper, flags.p);
            AB_LAUNCH_CHECK();
            AB_TRY(scan_exclusive_u32(flags.p, flags.p, nnz, total.p));
            uint32_t nextnz = 0;
            AB_CUDA(cudaMemcpyAsync(&nextnz, total.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream()));
            AB_CUDA(cudaStreamSynchronize(stream()));
            if (nextnz) {
                DevBuf<uint64_t> k0, k1; DevBuf<uint32_t> p0, p1, uf;
                AB_TRY(k0.alloc(nextnz)); AB_TRY(k1.alloc(nextnz)); AB_TRY(p0.alloc(nextnz)); AB_TRY(p1.alloc(nextnz));
                ext_compact_kernel<<<nb, 256, 0, stream()>>>(src->colind, nnz, ilower, iupper, flags.p, k0.p);
                AB_LAUNCH_CHECK();
                uint64_t* ks; uint32_t* ps;
                int nbits = 1; while (nbits < 40 && ((int64_t)1 << nbits) < N) ++nbits;
                AB_TRY(radix_sort_u64(k0.p, p0.p, k1.p, p1.p, nextnz, nbits, &ks, &ps));
                AB_TRY(uf.alloc(nextnz));
                unsigned nb2 = (unsigned)((nextnz + 255) / 256);
                uniq_flag_kernel<<<nb2, 256, 0, stream()>>>(ks, nextnz, uf.p);
                AB_LAUNCH_CHECK();
                AB_TRY(scan_exclusive_u32(uf.p, uf.p, nextnz, total.p));
                AB_CUDA(cudaMemcpyAsync(&next, total.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream()));
                AB_CUDA(cudaStreamSynchronize(stream()));
                AB_TRY(ext.alloc(next));
                uniq_emit_kernel<<<nb2, 256, 0, stream()>>>(ks, nextnz, uf.p, ext.p);
                AB_LAUNCH_CHECK();
            }
        }
        d->next = next;
        if (g_world == 1 && next) return fail(AB_EINVAL, "matrix has %u columns outside the only rank's rows", next);
        // ---- 3. re-indexed local block
        Csr* c = new Csr();
        d->loc = c;
        c->m = nloc; c->n = nloc + next; c->nnz = src->nnz; c->T = src->nnz;
        AB_TRY(dev_alloc((void**)&c->rowptr, (size_t)(nloc + 1) * sizeof(int32_t)));
        AB_TRY(dev_alloc((void**)&c->colind, (nnz + 8) * sizeof(int32_t)));
        AB_TRY(dev_alloc((void**)&c->values, (nnz + 8) * sizeof(double)));
        AB_CUDA(cudaMemcpyAsync(c->rowptr, src->rowptr, (size_t)(nloc + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, stream()));
        AB_CUDA(cudaMemcpyAsync(c->values, src->values, (nnz + 8) * sizeof(double), cudaMemcpyDeviceToDevice, stream()));
        AB_CUDA(cudaMemsetAsync(c->colind + nnz, 0, 8 * sizeof(int32_t), stream()));
        if (nnz) {
            remap_cols_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, stream()>>>(src->colind, nnz, ilower, iupper, nloc, ext.p, next, c->colind);
            AB_LAUNCH_CHECK();
        }
        AB_TRY(csr_finalize(c));
        AB_TRY(dev_alloc((void**)&d->xext, (size_t)(nloc + next + 2) * sizeof(double)));
        AB_CUDA(cudaMemsetAsync(d->xext, 0, (size_t)(nloc + next + 2) * sizeof(double), stream()));
        // ---- 4. who owns my halo columns (ext is sorted, owners are contiguous runs)
        d->send_cnt.assign(g_world, 0); d->send_off.assign(g_world, 0);
        d->recv_cnt.assign(g_world, 0); d->recv_off.assign(g_world, 0);
        if (g_world == 1) return AB_OK;
        std::vector<int64_t> hext(next);
        if (next) {
            AB_CUDA(cudaMemcpyAsync(hext.data(), ext.p, (size_t)next * sizeof(int64_t), cudaMemcpyDeviceToHost, stream()));
            AB_CUDA(cudaStreamSynchronize(stream()));
        }
        {
            size_t k = 0;
            for (int p = 0; p < g_world; ++p) {
                d->recv_off[p] = (int)k;
                const int64_t lo = ranges[2 * p], hi = ranges[2 * p + 1];
                size_t k0 = k;
                while (k < hext.size() && hext[k] >= lo && hext[k] <= hi) ++k;
                d->recv_cnt[p] = (int)(k - k0);
            }
            if (k != hext.size()) return fail(AB_EINVAL, "a halo column is owned by no rank (row ranges must tile [0,N))");
        }
        // ---- 5. tell each owner how many / which rows I need
        DevBuf<int32_t> dcnt_out, dcnt_in;
        AB_TRY(dcnt_out.alloc(g_world)); AB_TRY(dcnt_in.alloc(g_world));
        AB_CUDA(cudaMemcpyAsync(dcnt_out.p, d->recv_cnt.data(), (size_t)g_world * sizeof(int32_t), cudaMemcpyHostToDevice, stream()));
        AB_CUDA(cudaMemsetAsync(dcnt_in.p, 0, (size_t)g_world * sizeof(int32_t), stream()));
        AB_NCCL(g_nccl.GroupStart());
        for (int p = 0; p < g_world; ++p) {
            if (p == g_rank) continue;
            AB_NCCL(g_nccl.Send(dcnt_out.p + p, 1, ncclInt32, p, g_comm, stream()));
            AB_NCCL(g_nccl.Recv(dcnt_in.p + p, 1, ncclInt32, p, g_comm, stream()));
        }
        AB_NCCL(g_nccl.GroupEnd());
        AB_CUDA(cudaMemcpyAsync(d->send_cnt.data(), dcnt_in.p, (size_t)g_world * sizeof(int32_t), cudaMemcpyDeviceToHost, stream()));
        AB_CUDA(cudaStreamSynchronize(stream()));
        int64_t ns = 0;
        for (int p = 0; p < g_world; ++p) { d->send_off[p] = (int)ns; ns += d->send_cnt[p]; }
        d->nsend = ns;
        DevBuf<int64_t> gsend;
        AB_TRY(gsend.alloc((size_t)ns));
        AB_NCCL(g_nccl.GroupStart());
        for (int p = 0; p < g_world; ++p) {
            if (p == g_rank) continue;
            if (d->recv_cnt[p]) AB_NCCL(g_nccl.Send(ext.p + d->recv_off[p], (size_t)d->recv_cnt[p], ncclInt64, p, g_comm, stream()));
            if (d->send_cnt[p]) AB_NCCL(g_nccl.Recv(gsend.p + d->send_off[p], (size_t)d->send_cnt[p], ncclInt64, p, g_comm, stream()));
        }
        AB_NCCL(g_nccl.GroupEnd());
        AB_TRY(dev_alloc((void**)&d->send_idx, (size_t)(ns > 0 ? ns : 1) * sizeof(int32_t)));
        AB_TRY(dev_alloc((void**)&d->sendbuf, (size_t)(ns > 0 ? ns : 1) * sizeof(double)));
        if (ns) {
            DevBuf<int> err;
            AB_TRY(err.alloc(1));
            AB_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), stream()));
            to_local_idx_kernel<<<(unsigned)((ns + 255) / 256), 256, 0, stream()>>>(gsend.p, (size_t)ns, ilower, iupper, d->send_idx, err.p);
            AB_LAUNCH_CHECK();
            int herr = 0;
            AB_CUDA(cudaMemcpyAsync(&herr, err.p, sizeof(int), cudaMemcpyDeviceToHost, stream()));
            AB_CUDA(cudaStreamSynchronize(stream()));
            if (herr) return fail(AB_EINVAL, "a peer asked for a row this rank does not own");
        }
        AB_CUDA(cudaStreamSynchronize(stream()));
        return AB_OK;
    };
    int rc = body();
    if (rc != AB_OK) { delete d; return rc; }
    {
        std::lock_guard<std::mutex> lk2(g_dist_mu);
        *out_dh = g_next_dh++;
        g_dist[*out_dh] = d;
    }
    return AB_OK;
}

int ab_dist_destroy(int64_t dh) {
    AB_TRY(ensure_device());
    DistCsr* d = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_dist_mu);
        auto it = g_dist.find(dh);
        if (it == g_dist.end()) return fail(AB_EHANDLE, "unknown distributed handle %lld", (long long)dh);
        d = it->second;
        g_dist.erase(it);
    }
    delete d;
    return AB_OK;
}

int ab_dist_spmv(int64_t dh, const double* x_, double* y_) {
    AB_TRY(ensure_device());
    DistCsr* d = dist_get(dh);
    if (!d) return AB_EHANDLE;
    if (d->nloc && (!x_ || !y_)) return fail(AB_EINVAL, "null vector");
    std::lock_guard<std::mutex> lk(big_lock());
    Out<double> y;
    AB_TRY(y.bind(y_, (size_t)d->nloc));
    if (d->nloc) AB_CUDA(cudaMemcpyAsync(d->xext, x_, (size_t)d->nloc * sizeof(double), cudaMemcpyDefault, stream()));
    AB_TRY(halo_exchange(d));
    AB_TRY(launch_spmv(d->loc, d->xext, y.d, nullptr, nullptr, nullptr, nullptr, nullptr));
    AB_TRY(y.commit());
    return AB_OK;
}

static int dist_solve_common(int64_t dh, const double* f_, double* u_, double tol, int maxit, int* iters, double* relres) {
    AB_TRY(ensure_device());
    DistCsr* d = dist_get(dh);
    if (!d) return AB_EHANDLE;
    if (d->nloc && (!f_ || !u_)) return fail(AB_EINVAL, "null vector");
    std::lock_guard<std::mutex> lk(big_lock());
    In<double> f; Out<double> u;
    AB_TRY(f.bind(f_, (size_t)d->nloc)); AB_TRY(u.bind(u_, (size_t)d->nloc));
    DevBuf<double> xal;
    double* x = u.d;
    if (((uintptr_t)x & 15) != 0 || d->nloc == 0) { AB_TRY(xal.alloc((size_t)d->nloc + 2)); x = xal.p; }
    cudaEvent_t e0, e1;
    AB_CUDA(cudaEventCreate(&e0)); AB_CUDA(cudaEventCreate(&e1));
    AB_CUDA(cudaEventRecord(e0, stream()));
    SolveScal fin;
    int rc = dist_run_cg(d, f.d, x, tol, maxit, &fin);
    cudaEventRecord(e1, stream()); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    timer_add(AB_TIMER_SOLVE, ms * 1e-3);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (rc != AB_OK) return rc;
    if (x != u.d && d->nloc) AB_CUDA(cudaMemcpyAsync(u.d, x, (size_t)d->nloc * sizeof(double), cudaMemcpyDeviceToDevice, stream()));
    AB_TRY(u.commit());
    AB_CUDA(cudaStreamSynchronize(stream()));
    const double rel = fin.bb > 0 ? sqrt(fin.rr / fin.bb) : 0.0;
    if (iters) *iters = fin.iters;
    if (relres) *relres = rel;
    if (fin.breakdown) return fail(AB_EBREAKDOWN, "distributed CG breakdown after %d iterations (matrix not SPD?)", fin.iters);
    if (tol >= 0 && !fin.done) return fail(AB_ENOCONV, "distributed CG relres %.3e > tol %.3e after %d iterations", rel, tol, fin.iters);
    return AB_OK;
}

int ab_dist_solve(int64_t dh, const char* method, const double* f, double* u, double tol, int maxit, int* iters,
                  double* relres) {
    std::string m = method ? method : "CG";
    // the distributed path replaces Hypre BoomerAMG / GMRES (src/options.jl:54-58) with Jacobi-PCG
    if (!(m == "CG" || m == "cg" || m == "auto" || m == "BoomerAMG" || m == "GMRES"))
        return fail(AB_EINVAL, "unknown distributed solver '%s' (CG only)", m.c_str());
    if (tol <= 0) tol = opt("tol", 1e-10);   // Hypre tolerance in the reference: 1e-10 (MPITensor.h:66,91)
    if (maxit <= 0) maxit = (int)opt("maxit", 100000);
    return dist_solve_common(dh, f, u, tol, maxit, iters, relres);
}

int ab_dist_cg_fixed(int64_t dh, const double* f, double* u, int iters, double* relres) {
    if (iters < 0) return fail(AB_EINVAL, "iters < 0");
    return dist_solve_common(dh, f, u, -1.0, iters, nullptr, relres);
}

}  // extern "C"
