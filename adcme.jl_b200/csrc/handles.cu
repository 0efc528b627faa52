// handles.cu — process-global state: device binding, errors, options, timers, handle table.
// Mirrors the reference's "global map keyed by an integer handle" convention
// (SAMAP, deps/CustomOps/SparseAccumulate/SparseAccumulator.cpp:8).
#include "common.cuh"
#include <cstring>
#include <map>
#include <string>
#include <atomic>

namespace ab {

static thread_local char g_err[1024] = "";
static std::mutex g_mu;
static std::mutex g_big;
static bool g_bound        = false;
static int  g_device       = 0;
static int  g_sm_count     = 148;
static cudaStream_t g_stream = 0;
static std::atomic<long long> g_launches{0};
static std::map<std::string, double> g_opts;
static double g_timers[3] = {0, 0, 0};
static std::map<int64_t, Csr*> g_csr;
static int64_t g_next_h = 1;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

std::mutex& big_lock() { return g_big; }
cudaStream_t stream() { return g_stream; }
int sm_count() { return g_sm_count; }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static int bind_device(int device) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(AB_ECUDA, "no CUDA device visible (%s); libadcme_b200 has no CPU fallback",
                    e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
    }
    if (device < 0 || device >= ndev) return fail(AB_EINVAL, "device %d out of range [0,%d)", device, ndev);
    AB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    AB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(AB_ECUDA, "device %d is sm_%d%d; this library ships sm_100a code only", device,
                    prop.major, prop.minor);
    g_sm_count = prop.multiProcessorCount;
    g_device   = device;
    g_bound    = true;
    // keep freed blocks in the async pool instead of returning them to the OS
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    return AB_OK;
}

int ensure_device() {
    if (g_bound) {
        // a host thread other than the binder may call in (Julia tasks, TF inter-op pool)
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != g_device) cudaSetDevice(g_device);
        return AB_OK;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_bound) return AB_OK;
    int cur = 0;
    if (cudaGetDevice(&cur) != cudaSuccess) { cudaGetLastError(); cur = 0; }
    return bind_device(cur);
}

double opt(const char* key, double dflt) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_opts.find(key);
    return it == g_opts.end() ? dflt : it->second;
}
void timer_add(int which, double s) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (which >= 0 && which < 3) g_timers[which] += s;
}

int dev_alloc(void** p, size_t bytes) {
    *p = nullptr;
    cudaError_t e = cudaMallocAsync(p, bytes, g_stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(AB_ECUDA, "cudaMallocAsync(%zu bytes) -> %s", bytes, cudaGetErrorString(e));
    }
    return AB_OK;
}
int dev_free(void* p) {
    if (!p) return AB_OK;
    cudaError_t e = cudaFreeAsync(p, g_stream);
    if (e != cudaSuccess) { cudaGetLastError(); return AB_ECUDA; }
    return AB_OK;
}

bool is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// ------------------------------------------------------------- CSR table
Csr::~Csr() {
    dev_free(rowptr); dev_free(colind); dev_free(values);
    dev_free(perm); dev_free(segstart); dev_free(slot); dev_free(entry_row);
    dev_free(dinv); dev_free(tperm); dev_free(work); dev_free(scal);
    dev_free(partials); dev_free(ticket);
    dev_free(svalues); dev_free(sinv); dev_free(off8); dev_free(offtab); dev_free(code8); dev_free(codetab); dev_free(rowcode); dev_free(pat_ent); dev_free(pat_info);
    dev_free(rw_ent); dev_free(rw_tmask); dev_free(rw_mask); free(rw_plan);
    dev_free(rm_ent); dev_free(rm_items); free(rm_plan); dev_free(rm_rest); dev_free(rw_exclude);
    dev_free(zm_items); dev_free(zm_ent); dev_free(zm_rest); free(zm_plan); dev_free(zm_mask); dev_free(zm_rowmask); dev_free(zm_cta_first);
    dev_free(long_rows); dev_free(long_row_chunk0); dev_free(long_chunk_row); dev_free(long_chunk_start);
}

Csr* csr_get(int64_t h) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_csr.find(h);
    if (it == g_csr.end()) {
        set_error("unknown CSR handle %lld", (long long)h);
        return nullptr;
    }
    return it->second;
}
int64_t csr_put(Csr* c) {
    std::lock_guard<std::mutex> lk(g_mu);
    int64_t h = g_next_h++;
    g_csr[h]  = c;
    return h;
}
int csr_erase(int64_t h) {
    Csr* c = nullptr;
    int64_t th = 0;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_csr.find(h);
        if (it == g_csr.end()) return fail(AB_EHANDLE, "unknown CSR handle %lld", (long long)h);
        c = it->second;
        g_csr.erase(it);
        th = c->th;
    }
    delete c;
    if (th) csr_erase(th);
    return AB_OK;
}

}  // namespace ab

using namespace ab;

extern "C" {

int ab_init(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_bound && device == g_device) return AB_OK;
    return bind_device(device);
}

int ab_finalize(void) {
    std::vector<int64_t> hs;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        for (auto& kv : g_csr) hs.push_back(kv.first);
    }
    for (int64_t h : hs) {
        bool present;
        {
            std::lock_guard<std::mutex> lk(g_mu);
            present = g_csr.count(h) != 0;
        }
        if (present) csr_erase(h);
    }
    if (g_bound) cudaStreamSynchronize(g_stream);
    return AB_OK;
}

int ab_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* ab_last_error(void) { return g_err; }
const char* ab_version(void) { return "adcme_b200 0.1 (sm_100a)"; }

int ab_set_stream(void* s) {
    g_stream = (cudaStream_t)s;
    return AB_OK;
}
int ab_synchronize(void) {
    AB_TRY(ensure_device());
    AB_CUDA(cudaStreamSynchronize(g_stream));
    return AB_OK;
}
int ab_set_option(const char* key, double value) {
    if (!key) return fail(AB_EINVAL, "null option key");
    if (!strcmp(key, "l2_fetch_granularity")) {
        // device-wide hint (32 / 64 / 128 bytes): how much the L2 fetches from DRAM around a missing sector.  The random
        // 8-byte gathers of the assembly / gradient paths pay for every byte of it.
        AB_TRY(ensure_device());
        AB_CUDA(cudaDeviceSynchronize());
        AB_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value));
    }
    std::lock_guard<std::mutex> lk(g_mu);
    g_opts[key] = value;
    return AB_OK;
}
int ab_get_option(const char* key, double* value) {
    if (!key || !value) return fail(AB_EINVAL, "null argument");
    if (!strcmp(key, "l2_fetch_granularity")) {
        size_t v = 0;
        AB_TRY(ensure_device());
        AB_CUDA(cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity));
        *value = (double)v;
        return AB_OK;
    }
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_opts.find(key);
    if (it == g_opts.end()) return fail(AB_EINVAL, "option '%s' is not set", key);
    *value = it->second;
    return AB_OK;
}

int ab_timer_reset(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_timers[0] = g_timers[1] = g_timers[2] = 0;
    return AB_OK;
}
double ab_timer_get(int which) {
    std::lock_guard<std::mutex> lk(g_mu);
    return (which >= 0 && which < 3) ? g_timers[which] : -1.0;
}
int64_t ab_launch_count(void) { return (int64_t)g_launches.load(); }

int ab_csr_query(int64_t h, int64_t* m, int64_t* n, int64_t* nnz, int64_t* T) {
    Csr* c = csr_get(h);
    if (!c) return AB_EHANDLE;
    if (m) *m = c->m;
    if (n) *n = c->n;
    if (nnz) *nnz = c->nnz;
    if (T) *T = c->T;
    return AB_OK;
}
int ab_csr_destroy(int64_t h) {
    AB_TRY(ensure_device());
    return csr_erase(h);
}

}  // extern "C"
