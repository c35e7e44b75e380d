// C ABI of libpmf (include/pmf.h): handle, marshalling, epoch orchestration.
// The only execution engine behind this ABI is CUDA; there is no CPU path.
#include "../../include/pmf.h"
#include "pmf_internal.h"
#include "pmf_host.h"

#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

using namespace pmf;

static thread_local std::string g_last_error;

static int fail(pmf_model_s* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    g_last_error = buf;
    return code;
}

#define CU(h, expr)                                                                              \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            (h)->cuda_failed = true;                                                             \
            return fail((h), PMF_ERR_CUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e__), \
                        __FILE__, __LINE__, #expr);                                              \
        }                                                                                        \
    } while (0)

#define CHECK_H(h)                                                              \
    do {                                                                        \
        if (!(h)) return fail(nullptr, PMF_ERR_ARG, "null handle");             \
        if ((h)->cuda_failed) return PMF_ERR_CUDA;                              \
        cudaError_t e0__ = cudaSetDevice((h)->dims.device);                     \
        if (e0__ != cudaSuccess) {                                              \
            return fail((h), PMF_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e0__)); \
        }                                                                       \
    } while (0)

// Every host <-> device copy of the ABI runs on the handle's stream and completes before the call returns: the
// stream is non-blocking, so the legacy-stream cudaMemcpy / cudaMemset would not be ordered against its kernels
// (a pageable H2D cudaMemcpy may return before its DMA has landed).
static cudaError_t copy_sync(cudaStream_t s, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, s);
    return e != cudaSuccess ? e : cudaStreamSynchronize(s);
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

template <class T>
static cudaError_t dev_alloc(T** p, size_t n) {
    *p = nullptr;
    if (n == 0) return cudaSuccess;
    return cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T));
}
template <class T>
static void dev_free(T*& p) {
    if (p) cudaFree((void*)p);
    p = nullptr;
}

// The data buffer (and its permuted copy on the batch path) is by far the largest allocation of a handle, and
// cudaMalloc / cudaFree of a gigabyte cost 10-200 ms depending on the driver's mood.  Callers that create a
// handle per fit (mf_fit on a host-resident model) would pay that on every call, so a destroyed handle parks its
// big buffers here and the next pmf_create of the same size on the same device takes them over.  At most
// kPoolSlots buffers are parked; pmf_release_cached_memory() frees them.
namespace {
struct ParkedBuf { int dev; size_t bytes; void* p; };
constexpr size_t kPoolMinBytes = 64u << 20;
constexpr int kPoolSlots = 2;
std::mutex g_pool_mu;
std::vector<ParkedBuf> g_pool;
// the pinned control block of a handle (cudaMallocHost / cudaFreeHost page-lock and unlock: ~0.1-1 ms each)
std::mutex g_ctrl_pool_mu;
std::vector<void*> g_ctrl_pool;
}  // namespace

static cudaError_t ctrl_host_alloc(FitControl** p) {
    {
        std::lock_guard<std::mutex> lk(g_ctrl_pool_mu);
        if (!g_ctrl_pool.empty()) {
            *p = static_cast<FitControl*>(g_ctrl_pool.back());
            g_ctrl_pool.pop_back();
            std::memset(*p, 0, sizeof(FitControl));
            return cudaSuccess;
        }
    }
    // portable: one pinned block serves the handles of every device of the process
    cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(p), sizeof(FitControl), cudaHostAllocPortable);
    if (e == cudaSuccess) std::memset(*p, 0, sizeof(FitControl));
    return e;
}
static void ctrl_host_free(FitControl*& p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_ctrl_pool_mu);
    if (g_ctrl_pool.size() < 16) g_ctrl_pool.push_back(p);
    else cudaFreeHost(p);
    p = nullptr;
}

static cudaError_t big_alloc(float** p, size_t n_floats, int dev) {
    const size_t bytes = n_floats * sizeof(float);
    *p = nullptr;
    if (bytes == 0) return cudaSuccess;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        for (size_t k = 0; k < g_pool.size(); ++k)
            if (g_pool[k].dev == dev && g_pool[k].bytes == bytes) {
                *p = static_cast<float*>(g_pool[k].p);
                g_pool.erase(g_pool.begin() + k);
                return cudaSuccess;
            }
    }
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), bytes);
    if (e != cudaSuccess) {      // out of memory: give the parked buffers back and retry once
        pmf_release_cached_memory();
        cudaGetLastError();
        e = cudaMalloc(reinterpret_cast<void**>(p), bytes);
    }
    return e;
}
// the caller has synchronised the device: no work can still touch the buffer
static void big_free(float*& p, size_t n_floats, int dev) {
    if (!p) return;
    const size_t bytes = n_floats * sizeof(float);
    void* victim = p;
    p = nullptr;
    if (bytes >= kPoolMinBytes) {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        g_pool.push_back(ParkedBuf{dev, bytes, victim});
        victim = nullptr;
        if ((int)g_pool.size() > kPoolSlots) {
            victim = g_pool.front().p;
            g_pool.erase(g_pool.begin());
        }
    }
    if (victim) cudaFree(victim);
}

int pmf_release_cached_memory(void) {
    std::vector<ParkedBuf> drop;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        drop.swap(g_pool);
    }
    int cur = 0;
    cudaGetDevice(&cur);
    for (const ParkedBuf& b : drop) {
        cudaSetDevice(b.dev);
        cudaFree(b.p);
    }
    cudaSetDevice(cur);
    pmf::release_alloc_cache();         // the small blocks parked by the allocator (guard.cu)
    {
        std::lock_guard<std::mutex> lk(g_ctrl_pool_mu);
        for (void* q : g_ctrl_pool) cudaFreeHost(q);
        g_ctrl_pool.clear();
    }
    return PMF_OK;
}

// host [rows][w] (pitch w) -> device [rows][wd] (pitch wd)
static cudaError_t up2d(float* dst, int wd, const float* src, int w, int rows, cudaStream_t s) {
    return cudaMemcpy2DAsync(dst, (size_t)wd * 4, src, (size_t)w * 4, (size_t)w * 4, rows, cudaMemcpyHostToDevice, s);
}
static cudaError_t down2d(float* dst, int w, const float* src, int wd, int rows, cudaStream_t s) {
    return cudaMemcpy2DAsync(dst, (size_t)w * 4, src, (size_t)wd * 4, (size_t)w * 4, rows, cudaMemcpyDeviceToHost, s);
}

// ---- NCCL through dlopen ------------------------------------------------------------------------
// Only the five entry points of the exchange step; types restated from nccl.h (ABI-stable since 2.x).
struct Id128 { char bytes[128]; };   // ncclUniqueId (passed by value)
namespace {
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
};
NcclApi& nccl() {
    static NcclApi api;
    if (!api.lib) {
        api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) {
            api.GetUniqueId = reinterpret_cast<int (*)(void*)>(dlsym(api.lib, "ncclGetUniqueId"));
            api.CommInitRank = reinterpret_cast<int (*)(void**, int, Id128, int)>(dlsym(api.lib, "ncclCommInitRank"));
            api.AllReduce = reinterpret_cast<int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t)>(dlsym(api.lib, "ncclAllReduce"));
            api.CommDestroy = reinterpret_cast<int (*)(void*)>(dlsym(api.lib, "ncclCommDestroy"));
            api.GetErrorString = reinterpret_cast<const char* (*)(int)>(dlsym(api.lib, "ncclGetErrorString"));
            api.GroupStart = reinterpret_cast<int (*)()>(dlsym(api.lib, "ncclGroupStart"));
            api.GroupEnd = reinterpret_cast<int (*)()>(dlsym(api.lib, "ncclGroupEnd"));
        }
    }
    return api;
}
bool nccl_ok(const NcclApi& a) {
    return a.lib && a.GetUniqueId && a.CommInitRank && a.AllReduce && a.CommDestroy && a.GroupStart && a.GroupEnd;
}
constexpr int NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8, NCCL_SUM = 0;   // ncclDataType_t / ncclRedOp_t (nccl.h)
}  // namespace

extern "C" {

const char* pmf_version(void) { return "libpmf 0.1 (sm_100a)"; }

const char* pmf_last_error(pmf_handle h) { return h ? h->err.c_str() : g_last_error.c_str(); }

int pmf_create(const pmf_dims* d, pmf_handle* out) {
    if (!d || !out) return fail(nullptr, PMF_ERR_ARG, "null argument");
    if (d->M <= 0 || d->N <= 0 || d->K <= 0) return fail(nullptr, PMF_ERR_ARG, "M, N, K must be positive");
    if (d->K > 256) return fail(nullptr, PMF_ERR_ARG, "K > 256 is not supported");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, PMF_ERR_CUDA, "no CUDA device: libpmf has no CPU path (%s)", cudaGetErrorString(e));
    if (d->device < 0 || d->device >= ndev) return fail(nullptr, PMF_ERR_ARG, "bad device ordinal %d", d->device);
    pmf_model_s* h = new pmf_model_s();
    h->dims = *d;
    h->M = d->M; h->N = d->N; h->K = d->K;
    h->Kp = round_up(d->K, 8);
    h->lda = round_up(d->M, 32);
    h->Mp = round_up(d->M, 128);
    h->Np = round_up(d->N, 128);
    if ((e = cudaSetDevice(d->device)) != cudaSuccess) {
        delete h;
        return fail(nullptr, PMF_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    }
    // two attribute queries, not cudaGetDeviceProperties (which fills ~100 fields and costs milliseconds per handle)
    cudaDeviceGetAttribute(&h->n_sms, cudaDevAttrMultiProcessorCount, d->device);
    cudaDeviceGetAttribute(&h->cc_major, cudaDevAttrComputeCapabilityMajor, d->device);
    if (h->cc_major != 10) {
        // the library holds sm_100a code only: say so here instead of failing at the first kernel launch
        const int cc = h->cc_major;
        delete h;
        return fail(nullptr, PMF_ERR_CUDA, "device %d has compute capability %d.x: libpmf is built for sm_100a (B200) only", d->device, cc);
    }
    cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    h->stream = h->own_stream;
    cudaEventCreate(&h->ev0);
    cudaEventCreate(&h->ev1);
    const size_t xk = (size_t)h->Mp * h->Kp, yk = (size_t)h->Np * h->Kp;
    bool ok = true;
    ok &= big_alloc(&h->A, (size_t)h->N * h->lda, d->device) == cudaSuccess;
    ok &= dev_alloc(&h->X, xk) == cudaSuccess && dev_alloc(&h->dX, xk) == cudaSuccess && dev_alloc(&h->accX, xk) == cudaSuccess;
    ok &= dev_alloc(&h->Y, yk) == cudaSuccess && dev_alloc(&h->accY, yk) == cudaSuccess;
    ok &= dev_alloc(&h->weight, h->Np) == cudaSuccess && dev_alloc(&h->colinfo, h->Np) == cudaSuccess;
    ok &= dev_alloc(&h->scalars_base, 4 * SC_COUNT) == cudaSuccess && dev_alloc(&h->ctrl_base, 2) == cudaSuccess;
    h->scalars = h->scalars_base; h->ctrl = h->ctrl_base;
    ok &= dev_alloc(&h->thresholds, 4) == cudaSuccess && dev_alloc(&h->acc_thr, 2 * PMF_MAX_RANGES) == cudaSuccess;
    if (!ok) {
        pmf_destroy(h);
        return fail(nullptr, PMF_ERR_ALLOC, "device allocation failed (A needs %.2f GB)", (double)h->N * h->lda * 4e-9);
    }
    cudaMemsetAsync(h->X, 0, xk * 4, h->stream); cudaMemsetAsync(h->Y, 0, yk * 4, h->stream);
    cudaMemsetAsync(h->colinfo, 0, (size_t)h->Np * 4, h->stream);
    cudaMemsetAsync(h->thresholds, 0, 16, h->stream);
    cudaMemsetAsync(h->ctrl_base, 0, 2 * sizeof(FitControl), h->stream);
    cudaMemsetAsync(h->scalars_base, 0, 4 * SC_COUNT * sizeof(double), h->stream);
    std::vector<float> ones(h->Np, 1.f);
    copy_sync(h->stream, h->weight, ones.data(), (size_t)h->Np * 4, cudaMemcpyHostToDevice);
    if (ctrl_host_alloc(&h->ctrl_host) != cudaSuccess) { pmf_destroy(h); return fail(nullptr, PMF_ERR_ALLOC, "pinned host allocation failed"); }
    int rc = h->realloc_vectors(0);   // no batch views yet
    if (rc != 0) { pmf_destroy(h); return fail(nullptr, PMF_ERR_ALLOC, "device allocation failed"); }
    pmf_reset_opt_state(h, 1e-8f);
    if (cudaDeviceSynchronize() != cudaSuccess) { pmf_destroy(h); return fail(nullptr, PMF_ERR_CUDA, "init failed"); }
    *out = h;
    return PMF_OK;
}

int pmf_destroy(pmf_handle h) {
    if (!h) return PMF_OK;
    cudaSetDevice(h->dims.device);
    cudaDeviceSynchronize();
    big_free(h->A, (size_t)h->N * h->lda, h->dims.device);
    dev_free(h->X); dev_free(h->dX); dev_free(h->accX); dev_free(h->Y); dev_free(h->accY);
    dev_free(h->Xl); dev_free(h->Xh); dev_free(h->tc_cost_cum);
    big_free(h->wide.G, (size_t)h->N * h->lda, h->dims.device);
    dev_free(h->wide.Xh); dev_free(h->wide.Yh);
    { float* q = static_cast<float*>(h->wide.Xb); dev_free(q); q = static_cast<float*>(h->wide.Yb); dev_free(q); h->wide.Xb = h->wide.Yb = nullptr; }
    dev_free(h->weight); dev_free(h->colinfo); dev_free(h->thresholds); dev_free(h->acc_thr); dev_free(h->scalars_base); dev_free(h->ctrl_base);
    h->scalars = nullptr; h->ctrl = nullptr;
    dev_free(h->vp); dev_free(h->sg); dev_free(h->accvp); dev_free(h->regw); dev_free(h->regc);
    dev_free(h->bcol_off); dev_free(h->bcol_view); dev_free(h->bcol_nb); dev_free(h->batch_of_sample);
    h->free_tc_plan();
    dev_free(h->hist); dev_free(h->col_ssq); dev_free(h->col_cnt); dev_free(h->col_sqerr);
    for (int s = 0; s < 2; ++s) h->reg[s].free_all();
    if (h->comm) { nccl().CommDestroy(h->comm); h->comm = nullptr; }
    ctrl_host_free(h->ctrl_host);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    for (cudaEvent_t ev : h->prof_ev) cudaEventDestroy(ev);
    delete h;
    return PMF_OK;
}

int pmf_set_stream(pmf_handle h, void* s) {
    CHECK_H(h);
    h->stream = s ? (cudaStream_t)s : h->own_stream;
    return PMF_OK;
}

int pmf_set_data(pmf_handle h, const float* A) {
    CHECK_H(h);
    if (!A) return fail(h, PMF_ERR_ARG, "null data");
    CU(h, cudaMemsetAsync(h->A, 0xFF, (size_t)h->N * h->lda * 4, h->stream));   // NaN padding
    CU(h, up2d(h->A, h->lda, A, h->M, h->N, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    h->have_data = true;
    h->tcb_valid = false;      // A_tc is a permuted copy of A
    return PMF_OK;
}

int pmf_set_factors(pmf_handle h, const float* X, const float* Y) {
    CHECK_H(h);
    if (X) CU(h, up2d(h->X, h->Kp, X, h->K, h->M, h->stream));
    if (Y) CU(h, up2d(h->Y, h->Kp, Y, h->K, h->N, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    h->xsplit_valid = false;
    return PMF_OK;
}

int pmf_get_factors(pmf_handle h, float* X, float* Y) {
    CHECK_H(h);
    if (X) CU(h, down2d(X, h->K, h->X, h->Kp, h->M, h->stream));
    if (Y) CU(h, down2d(Y, h->K, h->Y, h->Kp, h->N, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return PMF_OK;
}

int pmf_set_noise(pmf_handle h, int32_t n_ranges, const int32_t* cs, const int32_t* ce, const int32_t* dist,
                  const float* thresholds, const float* weight) {
    CHECK_H(h);
    if (n_ranges <= 0 || !cs || !ce || !dist) return fail(h, PMF_ERR_ARG, "bad noise ranges");
    if (n_ranges > PMF_MAX_RANGES) return fail(h, PMF_ERR_ARG, "more than %d noise ranges", PMF_MAX_RANGES);
    std::vector<int32_t> ci(h->Np, 0);
    std::vector<char> seen(h->N, 0);
    for (int r = 0; r < n_ranges; ++r) {
        if (cs[r] < 0 || ce[r] > h->N || cs[r] > ce[r]) return fail(h, PMF_ERR_ARG, "noise range %d out of bounds", r);
        if (dist[r] < 0 || dist[r] > 5) return fail(h, PMF_ERR_ARG, "unknown distribution code %d", dist[r]);
        for (int j = cs[r]; j < ce[r]; ++j) {
            if (seen[j]) return fail(h, PMF_ERR_ARG, "noise ranges overlap at column %d", j);
            seen[j] = 1;
            ci[j] = dist[r] | (r << 8);
        }
    }
    for (int j = 0; j < h->N; ++j)
        if (!seen[j]) return fail(h, PMF_ERR_ARG, "column %d has no noise model", j);
    dev_free(h->thresholds);
    CU(h, dev_alloc(&h->thresholds, (size_t)4 * n_ranges));
    std::vector<float> th(4 * (size_t)n_ranges, 0.f);
    if (thresholds) std::memcpy(th.data(), thresholds, th.size() * 4);
    CU(h, copy_sync(h->stream, h->thresholds, th.data(), th.size() * 4, cudaMemcpyHostToDevice));
    CU(h, copy_sync(h->stream, h->colinfo, ci.data(), (size_t)h->Np * 4, cudaMemcpyHostToDevice));
    if (weight) CU(h, copy_sync(h->stream, h->weight, weight, (size_t)h->N * 4, cudaMemcpyHostToDevice));
    {
        // Relative cost of one tile of the tcgen05 data pass per 128-feature tile (measured at the C2 shape: all-normal
        // 0.326 ms, all-bernoulli 0.431 ms, all-poisson 0.341 ms; inside the C2 mix the per-CTA clocks put a bernoulli
        // tile at 1.4 and a poisson tile at 1.07 normal tiles), cumulated so that the kernel can cut the tile list into
        // ranges of equal COST, not equal count.
        static const int kCost[6] = {100, 146, 106, 200, 130, 200};
        const int n_jt = (h->N + 127) / 128;
        std::vector<int32_t> cum(n_jt + 1, 0);
        for (int jt = 0; jt < n_jt; ++jt) {
            // a tile with several noise models pays extra: the warps that hold both take both branches in turn
            int w = 0, n_present = 0;
            unsigned present = 0;
            for (int j = jt * 128; j < std::min(h->N, (jt + 1) * 128); ++j) present |= 1u << (ci[j] & 0xff);
            for (int d = 0; d < 6; ++d)
                if (present & (1u << d)) { w = std::max(w, kCost[d]); ++n_present; }
            w += 30 * (n_present - 1);
            cum[jt + 1] = cum[jt] + w;
        }
        // the batch-path plan depends on the noise models only through these costs: a call that merely refreshes
        // the column weights (every mf_fit! on a resident model) keeps it
        std::vector<int32_t> cost(n_jt);
        for (int jt = 0; jt < n_jt; ++jt) cost[jt] = cum[jt + 1] - cum[jt];
        if (cost != h->tile_cost_host) {
            h->tile_cost_host.swap(cost);
            h->tcb_valid = false;
        }
        dev_free(h->tc_cost_cum);
        CU(h, dev_alloc(&h->tc_cost_cum, (size_t)n_jt + 1));
        CU(h, copy_sync(h->stream, h->tc_cost_cum, cum.data(), ((size_t)n_jt + 1) * 4, cudaMemcpyHostToDevice));
    }
    h->n_ranges = n_ranges;
    h->has_ordinal = false;
    for (int r = 0; r < n_ranges; ++r) h->has_ordinal |= (dist[r] == 3 || dist[r] == 5) && ce[r] > cs[r];
    h->have_noise = true;
    return PMF_OK;
}

int pmf_get_thresholds(pmf_handle h, int32_t n_ranges, float* thresholds) {
    CHECK_H(h);
    if (!thresholds || n_ranges != h->n_ranges) return fail(h, PMF_ERR_ARG, "pmf_get_thresholds: %d ranges given, the model has %d", n_ranges, h->n_ranges);
    CU(h, cudaStreamSynchronize(h->stream));
    CU(h, copy_sync(h->stream, thresholds, h->thresholds, (size_t)4 * n_ranges * 4, cudaMemcpyDeviceToHost));
    return PMF_OK;
}

int pmf_set_col_params(pmf_handle h, const float* logsigma, const float* mu) {
    CHECK_H(h);
    if (logsigma) CU(h, copy_sync(h->stream, h->logsigma(), logsigma, (size_t)h->N * 4, cudaMemcpyHostToDevice));
    if (mu) CU(h, copy_sync(h->stream, h->mu(), mu, (size_t)h->N * 4, cudaMemcpyHostToDevice));
    return PMF_OK;
}

int pmf_get_col_params(pmf_handle h, float* logsigma, float* mu) {
    CHECK_H(h);
    CU(h, cudaStreamSynchronize(h->stream));
    if (logsigma) CU(h, copy_sync(h->stream, logsigma, h->logsigma(), (size_t)h->N * 4, cudaMemcpyDeviceToHost));
    if (mu) CU(h, copy_sync(h->stream, mu, h->mu(), (size_t)h->N * 4, cudaMemcpyDeviceToHost));
    return PMF_OK;
}

int pmf_set_batch_layout(pmf_handle h, int32_t n_views, const int32_t* cs, const int32_t* ce, const int32_t* nb,
                         const int32_t* bos) {
    CHECK_H(h);
    if (n_views < 0 || (n_views > 0 && (!cs || !ce || !nb || !bos))) return fail(h, PMF_ERR_ARG, "bad batch layout");
    // An unchanged layout is a no-op: the batch parameters, their AdaGrad accumulators, the layer penalties and the
    // tcgen05 batch plan all survive.  Callers re-send the structure before every fit (mf_fit! on a resident model,
    // reweight_col_losses!, the LR-halving restarts of mf_fit_adapt_lr!, src/fit.jl:46-75), where the reference keeps
    // parameters and optimiser state.
    if (h->layout_set && n_views == (int)h->views.size()) {
        bool same = true;
        for (int v = 0; v < n_views && same; ++v)
            same = h->views[v].col_start == cs[v] && h->views[v].col_stop == ce[v] && h->views[v].n_batches == nb[v];
        if (same && n_views > 0)
            same = h->bos_host.size() == (size_t)n_views * h->M &&
                   std::memcmp(h->bos_host.data(), bos, h->bos_host.size() * sizeof(int32_t)) == 0;
        if (same) return PMF_OK;
    }
    CU(h, cudaStreamSynchronize(h->stream));
    // keep logsigma / mu across the re-allocation
    std::vector<float> ls(h->N), mu(h->N);
    CU(h, copy_sync(h->stream, ls.data(), h->logsigma(), (size_t)h->N * 4, cudaMemcpyDeviceToHost));
    CU(h, copy_sync(h->stream, mu.data(), h->mu(), (size_t)h->N * 4, cudaMemcpyDeviceToHost));
    h->views.clear();
    int64_t off = 0;
    int prev_end = 0, nb_max = 0;
    for (int v = 0; v < n_views; ++v) {
        if (cs[v] < prev_end || ce[v] > h->N || cs[v] >= ce[v] || nb[v] <= 0)
            return fail(h, PMF_ERR_ARG, "batch view %d: bad column range or batch count", v);
        prev_end = ce[v];
        BatchView bv{cs[v], ce[v], nb[v], off};
        off += (int64_t)(ce[v] - cs[v]) * nb[v];
        nb_max = std::max(nb_max, nb[v]);
        h->views.push_back(bv);
    }
    if (off > (int64_t)INT32_MAX) return fail(h, PMF_ERR_ARG, "batch table too large");
    h->nb_max = nb_max;
    if (h->realloc_vectors((int)off) != 0) return fail(h, PMF_ERR_ALLOC, "device allocation failed");
    CU(h, copy_sync(h->stream, h->logsigma(), ls.data(), (size_t)h->N * 4, cudaMemcpyHostToDevice));
    CU(h, copy_sync(h->stream, h->mu(), mu.data(), (size_t)h->N * 4, cudaMemcpyHostToDevice));
    dev_free(h->bcol_off); dev_free(h->bcol_view); dev_free(h->bcol_nb); dev_free(h->batch_of_sample);
    h->free_tc_plan();
    if (n_views > 0) {
        std::vector<int32_t> coff(h->Np, -1), cview(h->Np, -1), cnb(h->Np, 0);
        for (int v = 0; v < n_views; ++v)
            for (int j = cs[v]; j < ce[v]; ++j) {
                coff[j] = (int32_t)(h->views[v].offset + (int64_t)(j - cs[v]) * nb[v]);
                cview[j] = v;
                cnb[j] = nb[v];
            }
        std::vector<int32_t> b((size_t)n_views * h->Mp, 0);
        for (int v = 0; v < n_views; ++v)
            for (int i = 0; i < h->M; ++i) {
                int32_t x = bos[(size_t)v * h->M + i];
                if (x < 0 || x >= nb[v]) return fail(h, PMF_ERR_ARG, "batch_of_sample[%d][%d]=%d out of range", v, i, x);
                b[(size_t)v * h->Mp + i] = x;
            }
        CU(h, dev_alloc(&h->bcol_off, h->Np)); CU(h, dev_alloc(&h->bcol_view, h->Np)); CU(h, dev_alloc(&h->bcol_nb, h->Np));
        CU(h, dev_alloc(&h->batch_of_sample, b.size()));
        CU(h, copy_sync(h->stream, h->bcol_off, coff.data(), (size_t)h->Np * 4, cudaMemcpyHostToDevice));
        CU(h, copy_sync(h->stream, h->bcol_view, cview.data(), (size_t)h->Np * 4, cudaMemcpyHostToDevice));
        CU(h, copy_sync(h->stream, h->bcol_nb, cnb.data(), (size_t)h->Np * 4, cudaMemcpyHostToDevice));
        CU(h, copy_sync(h->stream, h->batch_of_sample, b.data(), b.size() * 4, cudaMemcpyHostToDevice));
        h->bos_host.assign(bos, bos + (size_t)n_views * h->M);
    } else {
        h->bos_host.clear();
    }
    h->layout_set = true;
    return PMF_OK;
}

static int check_view(pmf_model_s* h, int v) {
    if (v < 0 || v >= (int)h->views.size()) return fail(h, PMF_ERR_ARG, "batch view %d does not exist", v);
    return 0;
}

int pmf_set_batch_values(pmf_handle h, int32_t v, const float* logdelta, const float* theta) {
    CHECK_H(h);
    if (check_view(h, v)) return PMF_ERR_ARG;
    const BatchView& bv = h->views[v];
    size_t n = (size_t)(bv.col_stop - bv.col_start) * bv.n_batches;
    if (logdelta) CU(h, copy_sync(h->stream, h->logdelta() + bv.offset, logdelta, n * 4, cudaMemcpyHostToDevice));
    if (theta) CU(h, copy_sync(h->stream, h->theta() + bv.offset, theta, n * 4, cudaMemcpyHostToDevice));
    return PMF_OK;
}

int pmf_get_batch_values(pmf_handle h, int32_t v, float* logdelta, float* theta) {
    CHECK_H(h);
    if (check_view(h, v)) return PMF_ERR_ARG;
    CU(h, cudaStreamSynchronize(h->stream));
    const BatchView& bv = h->views[v];
    size_t n = (size_t)(bv.col_stop - bv.col_start) * bv.n_batches;
    if (logdelta) CU(h, copy_sync(h->stream, logdelta, h->logdelta() + bv.offset, n * 4, cudaMemcpyDeviceToHost));
    if (theta) CU(h, copy_sync(h->stream, theta, h->theta() + bv.offset, n * 4, cudaMemcpyDeviceToHost));
    return PMF_OK;
}

int pmf_get_batch_grads(pmf_handle h, int32_t v, float* dlogdelta, float* dtheta) {
    CHECK_H(h);
    if (check_view(h, v)) return PMF_ERR_ARG;
    CU(h, cudaStreamSynchronize(h->stream));
    const BatchView& bv = h->views[v];
    size_t n = (size_t)(bv.col_stop - bv.col_start) * bv.n_batches;
    if (dlogdelta) CU(h, copy_sync(h->stream, dlogdelta, h->g_logdelta() + bv.offset, n * 4, cudaMemcpyDeviceToHost));
    if (dtheta) CU(h, copy_sync(h->stream, dtheta, h->g_theta() + bv.offset, n * 4, cudaMemcpyDeviceToHost));
    return PMF_OK;
}

int pmf_set_frozen(pmf_handle h, uint32_t layer_mask, uint32_t reg_mask) {
    CHECK_H(h);
    h->frozen_layers = layer_mask & 0xF;
    h->frozen_regs = reg_mask & 0xF;
    return PMF_OK;
}

// ---- regularisers ----------------------------------------------------------------------------
static int side_n(pmf_model_s* h, int which) { return which == 0 ? h->M : h->N; }

int pmf_clear_reg(pmf_handle h, int32_t which) {
    CHECK_H(h);
    if (which < 0 || which > 1) return fail(h, PMF_ERR_ARG, "which must be 0 (X) or 1 (Y)");
    CU(h, cudaStreamSynchronize(h->stream));
    h->reg[which].free_all();
    return PMF_OK;
}

static std::vector<float> pad_scale(const float* w, int K, int Kp, float p) {
    std::vector<float> o(Kp, 0.f);
    for (int k = 0; k < K; ++k) o[k] = p * w[k];
    return o;
}

int pmf_set_reg_l2(pmf_handle h, int32_t which, const float* w, float p) {
    CHECK_H(h);
    if (which < 0 || which > 1 || !w) return fail(h, PMF_ERR_ARG, "bad L2 regulariser arguments");
    SideReg& r = h->reg[which];
    dev_free(r.l2_w);
    auto o = pad_scale(w, h->K, h->Kp, p);
    CU(h, dev_alloc(&r.l2_w, h->Kp));
    CU(h, copy_sync(h->stream, r.l2_w, o.data(), (size_t)h->Kp * 4, cudaMemcpyHostToDevice));
    return PMF_OK;
}

int pmf_set_reg_group(pmf_handle h, int32_t which, int32_t ng, const int32_t* st, const int32_t* en, const float* w, float p) {
    CHECK_H(h);
    if (which < 0 || which > 1 || ng <= 0 || !st || !en || !w) return fail(h, PMF_ERR_ARG, "bad group regulariser arguments");
    const int n = side_n(h, which);
    std::vector<int32_t> gid(n, -1);
    for (int g = 0; g < ng; ++g) {
        if (st[g] < 0 || en[g] > n || st[g] > en[g]) return fail(h, PMF_ERR_ARG, "group %d range out of bounds", g);
        for (int i = st[g]; i < en[g]; ++i) gid[i] = g;
    }
    std::vector<float> gw((size_t)ng * h->Kp, 0.f);
    for (int g = 0; g < ng; ++g)
        for (int k = 0; k < h->K; ++k) gw[(size_t)g * h->Kp + k] = p * w[(size_t)g * h->K + k];
    SideReg& r = h->reg[which];
    dev_free(r.group_id); dev_free(r.group_w);
    CU(h, dev_alloc(&r.group_id, n)); CU(h, dev_alloc(&r.group_w, gw.size()));
    CU(h, copy_sync(h->stream, r.group_id, gid.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    CU(h, copy_sync(h->stream, r.group_w, gw.data(), gw.size() * 4, cudaMemcpyHostToDevice));
    return PMF_OK;
}

int pmf_set_reg_sel_l1(pmf_handle h, int32_t which, const uint8_t* idx, const float* w, float p) {
    CHECK_H(h);
    if (which < 0 || which > 1 || !idx || !w) return fail(h, PMF_ERR_ARG, "bad selective-L1 arguments");
    const int n = side_n(h, which);
    std::vector<uint8_t> m((size_t)n * h->Kp, 0);
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < h->K; ++k) m[(size_t)i * h->Kp + k] = idx[(size_t)i * h->K + k] ? 1 : 0;
    SideReg& r = h->reg[which];
    dev_free(r.l1_mask); dev_free(r.l1_w);
    auto o = pad_scale(w, h->K, h->Kp, p);
    CU(h, dev_alloc(&r.l1_mask, m.size())); CU(h, dev_alloc(&r.l1_w, h->Kp));
    CU(h, copy_sync(h->stream, r.l1_mask, m.data(), m.size(), cudaMemcpyHostToDevice));
    CU(h, copy_sync(h->stream, r.l1_w, o.data(), (size_t)h->Kp * 4, cudaMemcpyHostToDevice));
    return PMF_OK;
}

int pmf_set_reg_ard(pmf_handle h, int32_t which, int32_t nr, const int32_t* st, const int32_t* en, const float* alpha,
                    const float* beta) {
    CHECK_H(h);
    if (which < 0 || which > 1 || nr <= 0 || !st || !en || !alpha || !beta) return fail(h, PMF_ERR_ARG, "bad ARD arguments");
    const int n = side_n(h, which);
    // rows outside every range get no penalty: alpha = -0.5 makes (0.5 + alpha) vanish
    std::vector<float> a(n, -0.5f), b(n, 1.f);
    for (int r = 0; r < nr; ++r) {
        if (st[r] < 0 || en[r] > n || st[r] > en[r]) return fail(h, PMF_ERR_ARG, "ARD range %d out of bounds", r);
        for (int i = st[r]; i < en[r]; ++i) { a[i] = alpha[r]; b[i] = beta[r]; }
    }
    SideReg& r = h->reg[which];
    dev_free(r.ard_alpha); dev_free(r.ard_beta_row); dev_free(r.ard_beta_full);
    CU(h, dev_alloc(&r.ard_alpha, n)); CU(h, dev_alloc(&r.ard_beta_row, n));
    CU(h, copy_sync(h->stream, r.ard_alpha, a.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    CU(h, copy_sync(h->stream, r.ard_beta_row, b.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
    return PMF_OK;
}

int pmf_set_reg_fsard(pmf_handle h, int32_t which, const float* alpha, const float* beta) {
    CHECK_H(h);
    if (which < 0 || which > 1 || !alpha || !beta) return fail(h, PMF_ERR_ARG, "bad FSARD arguments");
    const int n = side_n(h, which);
    const int npad = which == 0 ? h->Mp : h->Np;
    SideReg& r = h->reg[which];
    dev_free(r.ard_alpha); dev_free(r.ard_beta_row); dev_free(r.ard_beta_full);
    CU(h, dev_alloc(&r.ard_alpha, n)); CU(h, dev_alloc(&r.ard_beta_full, (size_t)npad * h->Kp));
    CU(h, copy_sync(h->stream, r.ard_alpha, alpha, (size_t)n * 4, cudaMemcpyHostToDevice));
    std::vector<float> ones((size_t)npad * h->Kp, 1.f);
    CU(h, copy_sync(h->stream, r.ard_beta_full, ones.data(), ones.size() * 4, cudaMemcpyHostToDevice));
    CU(h, up2d(r.ard_beta_full, h->Kp, beta, h->K, n, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return PMF_OK;
}

int pmf_get_fsard_beta(pmf_handle h, float* beta) {
    CHECK_H(h);
    SideReg& r = h->reg[1];
    if (!r.ard_beta_full || !beta) return fail(h, PMF_ERR_STATE, "no FSARD regulariser on Y");
    CU(h, down2d(beta, h->K, r.ard_beta_full, h->Kp, h->N, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return PMF_OK;
}

// builds the device copy of K concatenated CSR matrices
static cudaError_t upload_csr(cudaStream_t s, DevCsr& d, int K, const std::vector<int64_t>& rp_base, const std::vector<int64_t>& nnz_base,
                              const int32_t* rowptr, int64_t rp_total, const int32_t* col, const float* val, int64_t nnz_total) {
    cudaError_t e;
    if ((e = dev_alloc(&d.rowptr, (size_t)rp_total)) != cudaSuccess) return e;
    if ((e = dev_alloc(&d.col, (size_t)std::max<int64_t>(nnz_total, 1))) != cudaSuccess) return e;
    if ((e = dev_alloc(&d.val, (size_t)std::max<int64_t>(nnz_total, 1))) != cudaSuccess) return e;
    if ((e = dev_alloc(&d.rowptr_base, K)) != cudaSuccess) return e;
    if ((e = dev_alloc(&d.nnz_base, K)) != cudaSuccess) return e;
    copy_sync(s, d.rowptr, rowptr, (size_t)rp_total * 4, cudaMemcpyHostToDevice);
    if (nnz_total > 0) {
        copy_sync(s, d.col, col, (size_t)nnz_total * 4, cudaMemcpyHostToDevice);
        copy_sync(s, d.val, val, (size_t)nnz_total * 4, cudaMemcpyHostToDevice);
    }
    copy_sync(s, d.rowptr_base, rp_base.data(), (size_t)K * 8, cudaMemcpyHostToDevice);
    return copy_sync(s, d.nnz_base, nnz_base.data(), (size_t)K * 8, cudaMemcpyHostToDevice);
}

int pmf_set_reg_network(pmf_handle h, int32_t which, const int32_t* nv, const int32_t* aa_rp, const int32_t* aa_c,
                        const float* aa_v, const int32_t* ab_rp, const int32_t* ab_c, const float* ab_v,
                        const int32_t* bb_rp, const int32_t* bb_c, const float* bb_v, const float* xv, float p,
                        float rtol, float atol, int32_t itmax) {
    CHECK_H(h);
    if (which < 0 || which > 1 || !nv || !aa_rp || !ab_rp || !bb_rp) return fail(h, PMF_ERR_ARG, "bad network regulariser arguments");
    const int n = side_n(h, which), K = h->K;
    SideReg& r = h->reg[which];
    r.net.free_all();
    DevNetwork& net = r.net;
    std::vector<int64_t> aa_rb(K), aa_nb(K), ab_rb(K), ab_nb(K), bb_rb(K), bb_nb(K), vb(K);
    int64_t aa_rt = 0, aa_nt = 0, ab_rt = 0, ab_nt = 0, bb_rt = 0, bb_nt = 0, vt = 0;
    for (int k = 0; k < K; ++k) {
        if (nv[k] < 0) return fail(h, PMF_ERR_ARG, "nv[%d] < 0", k);
        aa_rb[k] = aa_rt; aa_nb[k] = aa_nt; aa_nt += aa_rp[aa_rt + n]; aa_rt += n + 1;
        ab_rb[k] = ab_rt; ab_nb[k] = ab_nt; ab_nt += ab_rp[ab_rt + n]; ab_rt += n + 1;
        bb_rb[k] = bb_rt; bb_nb[k] = bb_nt; bb_nt += bb_rp[bb_rt + nv[k]]; bb_rt += nv[k] + 1;
        vb[k] = vt; vt += nv[k];
    }
    // AB^T per factor (CSR with nv rows) built on the host
    std::vector<int32_t> abt_rp((size_t)bb_rt, 0), abt_c((size_t)std::max<int64_t>(ab_nt, 1));
    std::vector<float> abt_v((size_t)std::max<int64_t>(ab_nt, 1));
    for (int k = 0; k < K; ++k) {
        const int32_t* rp = ab_rp + ab_rb[k];
        const int32_t* c = ab_c + ab_nb[k];
        const float* v = ab_v + ab_nb[k];
        int32_t* trp = abt_rp.data() + bb_rb[k];
        for (int j = 0; j < n; ++j)
            for (int e = rp[j]; e < rp[j + 1]; ++e) {
                if (c[e] < 0 || c[e] >= nv[k]) return fail(h, PMF_ERR_ARG, "AB column index out of range (factor %d)", k);
                trp[c[e] + 1]++;
            }
        for (int u = 0; u < nv[k]; ++u) trp[u + 1] += trp[u];
        std::vector<int32_t> fill(trp, trp + nv[k]);
        for (int j = 0; j < n; ++j)
            for (int e = rp[j]; e < rp[j + 1]; ++e) {
                int32_t pos = fill[c[e]]++;
                abt_c[ab_nb[k] + pos] = j;
                abt_v[ab_nb[k] + pos] = v[e];
            }
    }
    CU(h, upload_csr(h->stream, net.AA, K, aa_rb, aa_nb, aa_rp, aa_rt, aa_c, aa_v, aa_nt));
    CU(h, upload_csr(h->stream, net.AB, K, ab_rb, ab_nb, ab_rp, ab_rt, ab_c, ab_v, ab_nt));
    CU(h, upload_csr(h->stream, net.BB, K, bb_rb, bb_nb, bb_rp, bb_rt, bb_c, bb_v, bb_nt));
    CU(h, upload_csr(h->stream, net.ABt, K, bb_rb, ab_nb, abt_rp.data(), bb_rt, abt_c.data(), abt_v.data(), ab_nt));
    CU(h, dev_alloc(&net.nv, K)); CU(h, dev_alloc(&net.virt_base, K));
    CU(h, copy_sync(h->stream, net.nv, nv, (size_t)K * 4, cudaMemcpyHostToDevice));
    CU(h, copy_sync(h->stream, net.virt_base, vb.data(), (size_t)K * 8, cudaMemcpyHostToDevice));
    CU(h, dev_alloc(&net.u, (size_t)std::max<int64_t>(vt, 1))); CU(h, dev_alloc(&net.work, (size_t)std::max<int64_t>(4 * vt, 1)));
    CU(h, cudaMemsetAsync(net.u, 0, (size_t)std::max<int64_t>(vt, 1) * 4, h->stream));
    if (xv && vt > 0) CU(h, copy_sync(h->stream, net.u, xv, (size_t)vt * 4, cudaMemcpyHostToDevice));
    net.nv_total = vt;
    net.nv_max = 0;
    for (int k = 0; k < K; ++k) net.nv_max = std::max(net.nv_max, (int)nv[k]);
    net.ld = round_up(n, 32);
    CU(h, dev_alloc(&net.yt, (size_t)K * net.ld)); CU(h, dev_alloc(&net.gt, (size_t)K * net.ld));
    net.p = p;
    const float deps = std::sqrt(1.1920929e-7f);
    net.rtol = rtol > 0 ? rtol : deps;
    net.atol = atol > 0 ? atol : deps;
    net.itmax = itmax;
    net.present = true;
    return PMF_OK;
}

int pmf_get_network_virtual(pmf_handle h, int32_t which, float* xv) {
    CHECK_H(h);
    if (which < 0 || which > 1 || !h->reg[which].net.present) return fail(h, PMF_ERR_STATE, "no network regulariser");
    CU(h, cudaStreamSynchronize(h->stream));
    DevNetwork& net = h->reg[which].net;
    if (net.nv_total > 0) CU(h, copy_sync(h->stream, xv, net.u, (size_t)net.nv_total * 4, cudaMemcpyDeviceToHost));
    return PMF_OK;
}

int pmf_set_layer_reg_col(pmf_handle h, int32_t slot, const float* w, const float* c) {
    CHECK_H(h);
    if (slot != 1 && slot != 3) return fail(h, PMF_ERR_ARG, "column layer regulariser slot must be 1 or 3");
    size_t off = slot == 1 ? 0 : (size_t)h->Np;
    h->layer_reg_present[slot - 1] = (w != nullptr);
    if (!w) return PMF_OK;
    CU(h, copy_sync(h->stream, h->regw + off, w, (size_t)h->N * 4, cudaMemcpyHostToDevice));
    if (c) CU(h, copy_sync(h->stream, h->regc + off, c, (size_t)h->N * 4, cudaMemcpyHostToDevice));
    else CU(h, cudaMemsetAsync(h->regc + off, 0, (size_t)h->N * 4, h->stream));
    return PMF_OK;
}

int pmf_set_layer_reg_batch(pmf_handle h, int32_t slot, const float* w, const float* c) {
    CHECK_H(h);
    if (slot != 2 && slot != 4) return fail(h, PMF_ERR_ARG, "batch layer regulariser slot must be 2 or 4");
    h->layer_reg_present[slot - 1] = (w != nullptr);
    if (!w) return PMF_OK;
    if (h->nbp == 0) return fail(h, PMF_ERR_STATE, "no batch layout set");
    // expand per-(view, batch) weights / centres to one value per table entry
    std::vector<float> ew((size_t)h->nbp), ec((size_t)h->nbp, 0.f);
    size_t src = 0;
    for (const BatchView& bv : h->views) {
        for (int jl = 0; jl < bv.col_stop - bv.col_start; ++jl)
            for (int b = 0; b < bv.n_batches; ++b) {
                ew[bv.offset + (size_t)jl * bv.n_batches + b] = w[src + b];
                if (c) ec[bv.offset + (size_t)jl * bv.n_batches + b] = c[src + b];
            }
        src += bv.n_batches;
    }
    size_t off = 2 * (size_t)h->Np + (slot == 2 ? 0 : (size_t)h->nbp);
    CU(h, copy_sync(h->stream, h->regw + off, ew.data(), ew.size() * 4, cudaMemcpyHostToDevice));
    CU(h, copy_sync(h->stream, h->regc + off, ec.data(), ec.size() * 4, cudaMemcpyHostToDevice));
    return PMF_OK;
}

// ---- optimiser state -----------------------------------------------------------------------
__global__ static void fill_kernel(float* p, size_t n, float v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

int pmf_reset_opt_state(pmf_handle h, float eps) {
    CHECK_H(h);
    fill_kernel<<<296, 256, 0, h->stream>>>(h->accX, (size_t)h->Mp * h->Kp, eps);
    fill_kernel<<<296, 256, 0, h->stream>>>(h->accY, (size_t)h->Np * h->Kp, eps);
    fill_kernel<<<296, 256, 0, h->stream>>>(h->accvp, h->vp_len(), eps);
    fill_kernel<<<1, 128, 0, h->stream>>>(h->acc_thr, 2 * PMF_MAX_RANGES, eps);
    CU(h, cudaGetLastError());
    CU(h, cudaStreamSynchronize(h->stream));
    return PMF_OK;
}

static int opt_locate(pmf_model_s* h, int which, int view, float** base, int* rows, int* w, int* wd) {
    switch (which) {
    case 0: *base = h->accX; *rows = h->M; *w = h->K; *wd = h->Kp; return 0;
    case 1: *base = h->accY; *rows = h->N; *w = h->K; *wd = h->Kp; return 0;
    case 2: *base = h->accvp; *rows = 1; *w = h->N; *wd = h->N; return 0;
    case 3: *base = h->accvp + h->Np; *rows = 1; *w = h->N; *wd = h->N; return 0;
    case 4: case 5: {
        if (check_view(h, view)) return -1;
        const BatchView& bv = h->views[view];
        *base = h->accvp + 2 * (size_t)h->Np + (which == 5 ? (size_t)h->nbp : 0) + bv.offset;
        *rows = 1; *w = *wd = (bv.col_stop - bv.col_start) * bv.n_batches;
        return 0;
    }
    default: fail(h, PMF_ERR_ARG, "bad optimiser-state selector %d", which); return -1;
    }
}

int pmf_get_opt_state(pmf_handle h, int32_t which, int32_t view, float* acc) {
    CHECK_H(h);
    float* base; int rows, w, wd;
    if (opt_locate(h, which, view, &base, &rows, &w, &wd)) return PMF_ERR_ARG;
    CU(h, down2d(acc, w, base, wd, rows, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return PMF_OK;
}

int pmf_set_opt_state(pmf_handle h, int32_t which, int32_t view, const float* acc) {
    CHECK_H(h);
    float* base; int rows, w, wd;
    if (opt_locate(h, which, view, &base, &rows, &w, &wd)) return PMF_ERR_ARG;
    CU(h, up2d(base, wd, acc, w, rows, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return PMF_OK;
}

// ---- epoch orchestration ----------------------------------------------------------------------
void pmf_default_fit_opts(pmf_fit_opts* o) {
    std::memset(o, 0, sizeof *o);
    o->max_epochs = 1000; o->epoch = 1; o->lr = 1.0f; o->adagrad_eps = 1e-8f;
    o->rel_tol = 1e-5; o->abs_tol = 1e-5;       // src/fit.jl:934-935
    o->kernel = PMF_KERNEL_AUTO; o->precision = 0; o->check_every = 8; o->no_terminate = 0;
    o->update_noise_models = 1;                  // src/fit.jl:14: true in every call of the reference
    o->alternating = 0;                          // one pass, simultaneous step
}

static int ready(pmf_model_s* h) {
    if (!h->have_data) return fail(h, PMF_ERR_STATE, "pmf_set_data has not been called");
    if (!h->have_noise) return fail(h, PMF_ERR_STATE, "pmf_set_noise has not been called");
    return 0;
}

static void fill_data_params(pmf_model_s* h, DataPassParams& p, bool use_stop) {
    std::memset(&p, 0, sizeof p);
    p.M = h->M; p.N = h->N; p.Kp = h->Kp; p.lda = h->lda; p.Mp = h->Mp; p.Np = h->Np;
    p.A = h->A; p.X = h->X; p.Y = h->Y;
    p.logsigma = h->logsigma(); p.mu = h->mu(); p.weight = h->weight;
    p.colinfo = h->colinfo; p.thresholds = h->thresholds;
    p.tc_cost_cum = h->tc_cost_cum;
    p.n_batch_views = (int)h->views.size(); p.nb_max = h->nb_max;
    p.bcol_off = h->bcol_off; p.bcol_view = h->bcol_view; p.bcol_nb = h->bcol_nb;
    p.batch_of_sample = h->batch_of_sample;
    p.logdelta = h->logdelta(); p.theta = h->theta();
    p.dX = h->dX; p.dY = h->g_Y(); p.dlogsigma = h->g_logsigma(); p.dmu = h->g_mu();
    p.dlogdelta = h->g_logdelta(); p.dtheta = h->g_theta();
    p.scalars = h->scalars;
    p.stop_flag = use_stop ? &h->ctrl->stop : nullptr;
    p.sample_chunks = 0;
    p.ordinal_eps = 1e-10f; p.hinge_margin = 1.0f;
}

// zero gradients (unless the previous epoch's update pass already did) + data pass
// (+ X-side penalties when `x_reg_now`: they are rank-local sums and must precede the exchange step)
static int phase_begin(pmf_model_s* h, const pmf_fit_opts* o, bool use_stop, bool x_reg_now) {
    cudaStream_t s = h->stream;
    if (!h->grads_clean) {
        CU(h, cudaMemsetAsync(h->dX, 0, (size_t)h->Mp * h->Kp * 4, s));
        CU(h, cudaMemsetAsync(h->sg, 0, h->sg_len() * 4, s));
        CU(h, cudaMemsetAsync(h->scalars, 0, SC_COUNT * 8, s));
    }
    h->grads_clean = false;
    DataPassParams p;
    fill_data_params(h, p, use_stop);
    if (h->has_ordinal && (o == nullptr || o->update_noise_models)) p.dthr = h->g_thr();
    int kind = o ? o->kernel : PMF_KERNEL_AUTO;
    int rc = h->run_data_pass(p, kind, o ? o->precision : 0);
    if (rc != 0) return rc;
    if (x_reg_now) {
        const int* stop = use_stop ? &h->ctrl->stop : nullptr;
        if ((rc = h->run_network_reg(0, stop)) != 0) return rc;
        if ((rc = h->run_reg_multi(true, false, false, stop)) != 0) return rc;
    }
    return 0;
}

// remaining penalties (loss + pullbacks added in place into the gradient buffers) in one launch
static int phase_reg_shared(pmf_model_s* h, bool use_stop, bool with_x) {
    const int* stop = use_stop ? &h->ctrl->stop : nullptr;
    int rc;
    if (with_x && (rc = h->run_network_reg(0, stop)) != 0) return rc;
    if ((rc = h->run_network_reg(1, stop)) != 0) return rc;
    return h->run_reg_multi(with_x, true, true, stop);
}

// which: 0 = every enabled parameter, 1 = column side only (Y, layers, thresholds), 2 = row side only (X)
static int phase_update(pmf_model_s* h, const pmf_fit_opts* o, int which = 0) {
    return h->run_update_multi(o->update_X != 0 && which != 1, o->update_Y != 0 && which != 2,
                               o->update_col_layers != 0 && which != 2, o->update_noise_models != 0 && which != 2, o->lr,
                               o->adagrad_eps, &h->ctrl->stop);
}

int pmf_loss_grad(pmf_handle h, int32_t include_reg, pmf_losses* out, float* dX, float* dY, float* dls, float* dmu) {
    CHECK_H(h);
    if (ready(h)) return PMF_ERR_STATE;
    pmf_fit_opts o;
    pmf_default_fit_opts(&o);
    o.kernel = h->loss_grad_kernel; o.precision = h->loss_grad_precision;
    h->grads_clean = false;
    int rc = phase_begin(h, &o, false, include_reg != 0);
    if (rc != 0) return rc;
    if (include_reg && (rc = phase_reg_shared(h, false, false)) != 0) return rc;
    CU(h, cudaStreamSynchronize(h->stream));
    double sc[SC_COUNT];
    CU(h, copy_sync(h->stream, sc, h->scalars, sizeof sc, cudaMemcpyDeviceToHost));
    if (out) {
        out->data = sc[SC_DATA]; out->x_reg = sc[SC_XREG]; out->y_reg = sc[SC_YREG]; out->layer_reg = sc[SC_LAYERREG];
        out->total = sc[SC_DATA] + sc[SC_XREG] + sc[SC_YREG] + sc[SC_LAYERREG];
    }
    if (dX) CU(h, down2d(dX, h->K, h->dX, h->Kp, h->M, h->stream));
    if (dY) CU(h, down2d(dY, h->K, h->g_Y(), h->Kp, h->N, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    if (dls) CU(h, copy_sync(h->stream, dls, h->g_logsigma(), (size_t)h->N * 4, cudaMemcpyDeviceToHost));
    if (dmu) CU(h, copy_sync(h->stream, dmu, h->g_mu(), (size_t)h->N * 4, cudaMemcpyDeviceToHost));
    return PMF_OK;
}

int pmf_get_threshold_grads(pmf_handle h, int32_t n_ranges, float* dthr) {
    CHECK_H(h);
    if (!dthr || n_ranges != h->n_ranges) return fail(h, PMF_ERR_ARG, "pmf_get_threshold_grads: %d ranges given, the model has %d", n_ranges, h->n_ranges);
    CU(h, cudaStreamSynchronize(h->stream));
    CU(h, copy_sync(h->stream, dthr, h->g_thr(), (size_t)2 * n_ranges * 4, cudaMemcpyDeviceToHost));
    return PMF_OK;
}

int pmf_fit_start(pmf_handle h, const pmf_fit_opts* o) {
    CHECK_H(h);
    if (!o) return fail(h, PMF_ERR_ARG, "null options");
    if (ready(h)) return PMF_ERR_STATE;
    if (o->epoch < 1 || o->max_epochs < o->epoch - 1) return fail(h, PMF_ERR_ARG, "bad epoch range [%d, %d]", o->epoch, o->max_epochs);
    int cap = std::max(1, o->max_epochs - o->epoch + 1);
    if (cap > h->hist_cap) {
        dev_free(h->hist);
        CU(h, dev_alloc(&h->hist, (size_t)5 * cap));
        h->hist_cap = cap;
    }
    FitControl c;
    std::memset(&c, 0, sizeof c);
    c.term_code = PMF_TERM_MAX_EPOCHS;
    c.epochs = o->epoch;
    h->scalars = h->scalars_base; h->ctrl = h->ctrl_base;
    CU(h, cudaMemcpyAsync(h->ctrl, &c, sizeof c, cudaMemcpyHostToDevice, h->stream));
    h->cur_epoch = o->epoch;
    h->launches = 0;
    h->grads_clean = false;
    return PMF_OK;
}

int pmf_epoch_begin(pmf_handle h, const pmf_fit_opts* o) {
    CHECK_H(h);
    if (!o) return fail(h, PMF_ERR_ARG, "null options");
    return phase_begin(h, o, true, true);     // X-side penalties before the caller's exchange step
}

int pmf_epoch_end(pmf_handle h, const pmf_fit_opts* o) {
    CHECK_H(h);
    if (!o) return fail(h, PMF_ERR_ARG, "null options");
    if (o->alternating) return fail(h, PMF_ERR_ARG, "alternating steps are only available through pmf_fit");
    int rc = phase_reg_shared(h, true, false);
    if (rc != 0) return rc;
    CU(h, launch_control(h->ctrl, h->scalars, h->hist, h->hist_cap, h->cur_epoch, o->no_terminate ? -1 : o->max_epochs, o->rel_tol, o->abs_tol, h->stream));
    h->launches++;
    rc = phase_update(h, o);
    if (rc != 0) return rc;
    h->cur_epoch++;
    return PMF_OK;
}

int pmf_fit_poll(pmf_handle h, pmf_history* out, int32_t* stopped) {
    CHECK_H(h);
    CU(h, cudaMemcpyAsync(h->ctrl_host, h->ctrl, sizeof(FitControl), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    const FitControl& c = *h->ctrl_host;
    if (stopped) *stopped = c.stop;
    if (out) {
        out->term_code = c.stop ? c.term_code : PMF_TERM_MAX_EPOCHS;
        out->epochs = c.epochs;
        int n = std::min(c.n_recorded, out->capacity);
        n = std::min(n, h->hist_cap);
        out->n_recorded = n;
        if (n > 0) {
            std::vector<double> hst((size_t)5 * n);
            CU(h, copy_sync(h->stream, hst.data(), h->hist, hst.size() * 8, cudaMemcpyDeviceToHost));
            for (int i = 0; i < n; ++i) {
                if (out->loss_total) out->loss_total[i] = hst[5 * i + 0];
                if (out->loss_data) out->loss_data[i] = hst[5 * i + 1];
                if (out->loss_x_reg) out->loss_x_reg[i] = hst[5 * i + 2];
                if (out->loss_y_reg) out->loss_y_reg[i] = hst[5 * i + 3];
                if (out->loss_layer_reg) out->loss_layer_reg[i] = hst[5 * i + 4];
            }
        }
        out->kernel_launches = h->launches;
    }
    return PMF_OK;
}

// The epoch loop with one fused pass per epoch after the data pass (reg_update.cu: fused_epoch_kernel).  Loss scalars
// rotate through a ring of three buffers (this epoch / next epoch, whose penalty values this epoch's pass produces at
// the updated parameters / the one after, cleared meanwhile), the control state through two.
static int fit_loop_fused(pmf_model_s* h, const pmf_fit_opts* o) {
    cudaStream_t s = h->stream;
    const bool sharded = h->comm != nullptr && h->comm_ranks > 1;
    const int check = o->check_every > 0 ? o->check_every : 8;
    int rc;
    CU(h, cudaMemsetAsync(h->scalars_base, 0, 4 * SC_COUNT * sizeof(double), s));
    if (!h->grads_clean) {
        CU(h, cudaMemsetAsync(h->dX, 0, (size_t)h->Mp * h->Kp * 4, s));
        CU(h, cudaMemsetAsync(h->sg, 0, h->sg_len() * 4, s));
    }
    h->scalars = h->scalars_base; h->ctrl = h->ctrl_base;
    if ((rc = h->run_penalty_values(o)) != 0) return rc;           // penalties of the first epoch
    int since = 0;
    for (int e = o->epoch; e <= o->max_epochs; ++e) {
        const int i = e - o->epoch;
        FitControl* cin = h->ctrl_base + (i & 1);
        FitControl* cout = h->ctrl_base + ((i + 1) & 1);
        h->scalars = h->scalars_base + (size_t)(i % 3) * SC_COUNT;
        h->ctrl = cin;
        h->grads_clean = true;          // cleared above / by the previous epoch's pass; the scalars ring is managed here
        if ((rc = phase_begin(h, o, true, false)) != 0) return rc;
        if ((rc = h->run_network_reg(0, &cin->stop)) != 0) return rc;      // rank-local: before the exchange
        if (sharded && (rc = h->exchange_gradients()) != 0) return rc;
        if ((rc = h->run_network_reg(1, &cin->stop)) != 0) return rc;
        FusedControl fc;
        fc.cin = cin; fc.cout = cout; fc.sc = h->scalars;
        fc.sc_zero = h->scalars_base + (size_t)((i + 2) % 3) * SC_COUNT;
        fc.hist = h->hist; fc.hist_cap = h->hist_cap; fc.epoch = h->cur_epoch;
        fc.max_epochs = o->no_terminate ? -1 : o->max_epochs; fc.rel_tol = o->rel_tol; fc.abs_tol = o->abs_tol;
        h->scalars = h->scalars_base + (size_t)((i + 1) % 3) * SC_COUNT;    // the pass adds the NEXT epoch's penalty values
        if ((rc = h->run_fused_epoch(o, fc)) != 0) return rc;
        h->ctrl = cout;
        h->cur_epoch++;
        if (++since >= check && e < o->max_epochs) {
            since = 0;
            CU(h, cudaMemcpyAsync(h->ctrl_host, cout, sizeof(FitControl), cudaMemcpyDeviceToHost, s));
            CU(h, cudaStreamSynchronize(s));
            if (h->ctrl_host->stop) break;
        }
    }
    h->scalars = h->scalars_base;
    return 0;
}

int pmf_fit(pmf_handle h, const pmf_fit_opts* o, pmf_history* out) {
    CHECK_H(h);
    int rc = pmf_fit_start(h, o);
    if (rc != 0) return rc;
    const int check = o->check_every > 0 ? o->check_every : 8;
    CU(h, cudaEventRecord(h->ev0, h->stream));
    int since = 0;
    if (!o->alternating) {
        if ((rc = fit_loop_fused(h, o)) != 0) return rc;
    } else
    for (int e = o->epoch; e <= o->max_epochs; ++e) {
        // Sharded: X-side penalties are rank-local sums and go before the exchange.  Single GPU: every
        // penalty of the epoch is one launch, after the data pass.
        const bool sharded = h->comm != nullptr && h->comm_ranks > 1;
        if ((rc = phase_begin(h, o, true, sharded)) != 0) return rc;
        if (sharded && (rc = h->exchange_gradients()) != 0) return rc;
        if ((rc = phase_reg_shared(h, true, !sharded)) != 0) return rc;
        CU(h, launch_control(h->ctrl, h->scalars, h->hist, h->hist_cap, h->cur_epoch, o->no_terminate ? -1 : o->max_epochs, o->rel_tol, o->abs_tol, h->stream));
        h->launches++;
        {
            // SURVEY App. D1, the other reading of MF.fit!: column-side step from this pass, then the row-side step from
            // a SECOND pass at the new column-side parameters (no history record, no termination test of its own; a
            // stop raised above makes every kernel of the second half return at once)
            if ((rc = phase_update(h, o, 1)) != 0) return rc;
            if (o->update_X) {
                if ((rc = phase_begin(h, o, true, sharded)) != 0) return rc;
                if (!sharded && (rc = h->run_network_reg(0, &h->ctrl->stop)) != 0) return rc;
                if (!sharded && (rc = h->run_reg_multi(true, false, false, &h->ctrl->stop)) != 0) return rc;
                if ((rc = phase_update(h, o, 2)) != 0) return rc;
            }
        }
        h->cur_epoch++;
        if (++since >= check && e < o->max_epochs) {
            since = 0;
            CU(h, cudaMemcpyAsync(h->ctrl_host, h->ctrl, sizeof(FitControl), cudaMemcpyDeviceToHost, h->stream));
            CU(h, cudaStreamSynchronize(h->stream));
            if (h->ctrl_host->stop) break;
        }
    }
    CU(h, cudaEventRecord(h->ev1, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    float ms = 0.f;
    CU(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    pmf_history tmp;
    std::memset(&tmp, 0, sizeof tmp);
    pmf_history* dst = out ? out : &tmp;
    rc = pmf_fit_poll(h, dst, nullptr);
    dst->device_ms = ms;
    return rc;
}

int pmf_comm_unique_id(uint8_t id_out[128]) {
    NcclApi& a = nccl();
    if (!nccl_ok(a)) return fail(nullptr, PMF_ERR_STATE, "libnccl.so.2 could not be loaded: %s", dlerror());
    if (!id_out) return fail(nullptr, PMF_ERR_ARG, "null id buffer");
    int rc = a.GetUniqueId(id_out);
    if (rc != 0) return fail(nullptr, PMF_ERR_CUDA, "ncclGetUniqueId: %s", a.GetErrorString ? a.GetErrorString(rc) : "?");
    return PMF_OK;
}

int pmf_comm_init_rank(pmf_handle h, int32_t n_ranks, int32_t rank, const uint8_t id[128]) {
    CHECK_H(h);
    NcclApi& a = nccl();
    if (!nccl_ok(a)) return fail(h, PMF_ERR_STATE, "libnccl.so.2 could not be loaded");
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks || !id) return fail(h, PMF_ERR_ARG, "bad communicator arguments");
    if (h->comm) pmf_comm_destroy(h);
    Id128 uid;
    std::memcpy(uid.bytes, id, 128);
    void* comm = nullptr;
    int rc = a.CommInitRank(&comm, n_ranks, uid, rank);
    if (rc != 0) return fail(h, PMF_ERR_CUDA, "ncclCommInitRank: %s", a.GetErrorString ? a.GetErrorString(rc) : "?");
    h->comm = comm;
    h->comm_ranks = n_ranks;
    return PMF_OK;
}

int pmf_comm_destroy(pmf_handle h) {
    CHECK_H(h);
    if (h->comm) {
        cudaStreamSynchronize(h->stream);
        nccl().CommDestroy(h->comm);
        h->comm = nullptr;
        h->comm_ranks = 1;
    }
    return PMF_OK;
}

int pmf_model_s::exchange_gradients() {
    if (!comm || comm_ranks <= 1) return 0;
    NcclApi& a = nccl();
    int rc = a.GroupStart();
    if (rc == 0) rc = a.AllReduce(sg, sg, sg_len(), NCCL_FLOAT32, NCCL_SUM, comm, stream);
    // SC_DATA and SC_XREG are rank-local partial sums; the other scalars are replicated and must not be summed
    if (rc == 0) rc = a.AllReduce(scalars, scalars, 2, NCCL_FLOAT64, NCCL_SUM, comm, stream);
    int rc2 = a.GroupEnd();
    if (rc == 0) rc = rc2;
    if (rc != 0) { cuda_failed = true; return fail(this, PMF_ERR_CUDA, "ncclAllReduce: %s", a.GetErrorString ? a.GetErrorString(rc) : "?"); }
    launches += 2;
    return 0;
}

int pmf_shared_grad_buffer(pmf_handle h, void** p, int64_t* n) {
    CHECK_H(h);
    if (p) *p = h->sg;
    if (n) *n = (int64_t)h->sg_len();
    return PMF_OK;
}

int pmf_shared_scalar_buffer(pmf_handle h, void** p, int64_t* n) {
    CHECK_H(h);
    if (p) *p = h->scalars;
    if (n) *n = 2;   // SC_DATA, SC_XREG are rank-local partial sums; the rest is replicated
    return PMF_OK;
}

// One streaming pass over the data with the current parameters (the FP32 tile kernel with a statistics
// epilogue): per column sum (dl/dz)^2, count of finite entries and squared error in link space; per
// (batch, column) of every batched view the count and the squared error (left in the batch-gradient tables).
static int run_stats_pass(pmf_model_s* h) {
    if (ready(h)) return PMF_ERR_STATE;
    if (!h->col_ssq) {
        CU(h, dev_alloc(&h->col_ssq, h->Np)); CU(h, dev_alloc(&h->col_cnt, h->Np)); CU(h, dev_alloc(&h->col_sqerr, h->Np));
    }
    CU(h, cudaMemsetAsync(h->col_ssq, 0, (size_t)h->Np * 4, h->stream));
    CU(h, cudaMemsetAsync(h->col_cnt, 0, (size_t)h->Np * 4, h->stream));
    CU(h, cudaMemsetAsync(h->col_sqerr, 0, (size_t)h->Np * 4, h->stream));
    if (h->nbp > 0) {
        CU(h, cudaMemsetAsync(h->g_logdelta(), 0, (size_t)h->nbp * 4, h->stream));
        CU(h, cudaMemsetAsync(h->g_theta(), 0, (size_t)h->nbp * 4, h->stream));
    }
    h->grads_clean = false;
    DataPassParams p;
    fill_data_params(h, p, false);
    p.col_ssq = h->col_ssq; p.col_cnt = h->col_cnt; p.col_sqerr = h->col_sqerr;
    size_t smem = sizeof(float) * ((size_t)128 * (h->Kp + 4) + 64 * 68 + 2 * (size_t)64 * h->nb_max);
    if (smem > 227 * 1024) return fail(h, PMF_ERR_ARG, "too many batches per view (%d) for K=%d", h->nb_max, h->K);
    CU(h, launch_data_pass_ffma(p, h->stream, h->n_sms));
    CU(h, cudaStreamSynchronize(h->stream));
    return PMF_OK;
}

int pmf_column_stats(pmf_handle h, float* ssq, float* nonnan) {
    CHECK_H(h);
    int rc = run_stats_pass(h);
    if (rc != 0) return rc;
    if (ssq) CU(h, copy_sync(h->stream, ssq, h->col_ssq, (size_t)h->N * 4, cudaMemcpyDeviceToHost));
    if (nonnan) CU(h, copy_sync(h->stream, nonnan, h->col_cnt, (size_t)h->N * 4, cudaMemcpyDeviceToHost));
    return PMF_OK;
}

int pmf_link_col_sqerr(pmf_handle h, float* sqerr, float* nonnan) {
    CHECK_H(h);
    int rc = run_stats_pass(h);
    if (rc != 0) return rc;
    if (sqerr) CU(h, copy_sync(h->stream, sqerr, h->col_sqerr, (size_t)h->N * 4, cudaMemcpyDeviceToHost));
    if (nonnan) CU(h, copy_sync(h->stream, nonnan, h->col_cnt, (size_t)h->N * 4, cudaMemcpyDeviceToHost));
    return PMF_OK;
}

int pmf_batch_stats(pmf_handle h, int32_t n_views, float* const* count, float* const* sqerr) {
    CHECK_H(h);
    if (n_views != (int32_t)h->views.size()) return fail(h, PMF_ERR_ARG, "pmf_batch_stats: %d views given, the model has %d", n_views, (int)h->views.size());
    int rc = run_stats_pass(h);
    if (rc != 0) return rc;
    for (int v = 0; v < n_views; ++v) {
        const BatchView& bv = h->views[v];
        size_t n = (size_t)(bv.col_stop - bv.col_start) * bv.n_batches;
        if (count && count[v]) CU(h, copy_sync(h->stream, count[v], h->g_theta() + bv.offset, n * 4, cudaMemcpyDeviceToHost));
        if (sqerr && sqerr[v]) CU(h, copy_sync(h->stream, sqerr[v], h->g_logdelta() + bv.offset, n * 4, cudaMemcpyDeviceToHost));
    }
    return PMF_OK;
}

/* test / bench hook (not part of the reference-facing surface): which kernel and precision
 * pmf_loss_grad uses. */
int pmf_set_loss_grad_kernel(pmf_handle h, int32_t kernel, int32_t precision) {
    CHECK_H(h);
    h->loss_grad_kernel = kernel;
    h->loss_grad_precision = precision;
    return PMF_OK;
}

int pmf_set_profiling(pmf_handle h, int32_t enable) {
    CHECK_H(h);
    h->profiling = enable != 0;
    h->prof_used = 0;
    return PMF_OK;
}

int pmf_get_profile(pmf_handle h, int32_t* n, float* mean_ms, float* min_ms) {
    CHECK_H(h);
    CU(h, cudaStreamSynchronize(h->stream));
    double sum = 0.0;
    float mn = 1e30f;
    int cnt = 0;
    for (size_t i = 0; i + 1 < h->prof_used; i += 2) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->prof_ev[i], h->prof_ev[i + 1]) == cudaSuccess) {
            sum += ms; mn = ms < mn ? ms : mn; cnt++;
        }
    }
    if (n) *n = cnt;
    if (mean_ms) *mean_ms = cnt ? (float)(sum / cnt) : 0.f;
    if (min_ms) *min_ms = cnt ? mn : 0.f;
    return PMF_OK;
}

}  // extern "C"

// ---- pmf_model_s members -------------------------------------------------------------------
int pmf_model_s::realloc_vectors(int new_nbp) {
    dev_free(vp); dev_free(sg); dev_free(accvp); dev_free(regw); dev_free(regc);
    nbp = new_nbp;
    const size_t nv = vp_len(), ng = sg_len();
    if (dev_alloc(&vp, nv) != cudaSuccess || dev_alloc(&sg, ng) != cudaSuccess || dev_alloc(&accvp, nv) != cudaSuccess ||
        dev_alloc(&regw, nv) != cudaSuccess || dev_alloc(&regc, nv) != cudaSuccess)
        return -1;
    cudaMemsetAsync(vp, 0, nv * 4, stream); cudaMemsetAsync(sg, 0, ng * 4, stream); cudaMemsetAsync(regw, 0, nv * 4, stream); cudaMemsetAsync(regc, 0, nv * 4, stream);
    fill_kernel<<<64, 256>>>(accvp, nv, 1e-8f);
    return cudaDeviceSynchronize() == cudaSuccess ? 0 : -1;
}

// ---- tcgen05 data pass with batch layers: sample orders and passes (TcBatchDev) -----------------------
void pmf_model_s::free_tc_plan() {
    if (tcb.A_tc) {
        cudaStreamSynchronize(stream);       // the buffer is parked for reuse, not freed: nothing may still read it
        float* a = const_cast<float*>(tcb.A_tc);
        big_free(a, (size_t)tcb.n_pass * 128 * tcb.n_pos, dims.device);
    }
    for (void* q : tcb_allocs) cudaFree(q);
    tcb_allocs.clear();
    tcb = pmf::TcBatchDev{};
    tcb_valid = false;
}

// Host-side planning of the batch path (pure bookkeeping, no device work): sample orders, passes, chunk batches.
struct TcBatchPlanHost {
    int n_orders = 1, n_pass = 0, n_pos = 0, n_used = 0;
    std::vector<std::vector<int32_t>> perms;         // [n_orders][n_pos] position -> sample, -1 = padding
    std::vector<int32_t> view_order, pass_feat0, pass_order, cost_cum;
    std::vector<uint16_t> boc;                       // [n_views][n_pos / 16]
};

static void plan_tc_batches(int M, int Mp, int N, const std::vector<BatchView>& views, const std::vector<int32_t>& bos_host,
                            const std::vector<int32_t>& tile_cost_host, TcBatchPlanHost& out) {
    const int V = (int)views.size();
    const int n_jt = (N + 127) / 128;
    // 1. sample orders (position -> sample, -1 = padding).  A view keeps the identity order when every 16-sample
    //    chunk lies in one batch as given; otherwise its samples are stably sorted by batch id and every batch is
    //    padded to a multiple of 16 positions.  Equal orders are shared between views.
    std::vector<std::vector<int32_t>> perms(1, std::vector<int32_t>(M));
    for (int p = 0; p < M; ++p) perms[0][p] = p;
    std::vector<int32_t> view_order(V, 0);
    for (int v = 0; v < V; ++v) {
        const int32_t* b = bos_host.data() + (size_t)v * M;
        bool uniform = true;
        for (int p = 1; p < M && uniform; ++p) uniform = (p & 15) == 0 || b[p] == b[p - 1];
        if (uniform) continue;
        std::vector<int32_t> idx(M);
        for (int p = 0; p < M; ++p) idx[p] = p;
        std::stable_sort(idx.begin(), idx.end(), [&](int32_t x, int32_t y) { return b[x] < b[y]; });
        std::vector<int32_t> perm;
        perm.reserve((size_t)M + 16 * (size_t)views[v].n_batches);
        for (int k = 0; k < M; ++k) {
            if (k > 0 && b[idx[k]] != b[idx[k - 1]])
                while (perm.size() & 15) perm.push_back(-1);
            perm.push_back(idx[k]);
        }
        int o = -1;
        for (int k = 1; k < (int)perms.size(); ++k)
            if (perms[k] == perm) { o = k; break; }
        if (o < 0) { perms.push_back(std::move(perm)); o = (int)perms.size() - 1; }
        view_order[v] = o;
    }
    const int n_orders = (int)perms.size();
    size_t n_used = 0;
    for (const auto& pm : perms) n_used = std::max(n_used, pm.size());
    const int n_pos = n_orders == 1 ? Mp : round_up((int)n_used, 128);
    for (auto& pm : perms) pm.resize(n_pos, -1);

    // 2. passes: every feature tile once per order its batched columns need (columns without batch layers
    //    ride with the tile's first pass)
    std::vector<int32_t> pass_feat0, pass_order, cost_cum(1, 0);
    for (int jt = 0; jt < n_jt; ++jt) {
        std::vector<int> os;
        for (int v = 0; v < V; ++v)
            if (views[v].col_start < std::min(N, 128 * jt + 128) && views[v].col_stop > 128 * jt &&
                std::find(os.begin(), os.end(), view_order[v]) == os.end())
                os.push_back(view_order[v]);
        if (os.empty()) os.push_back(0);
        for (int o : os) {
            pass_feat0.push_back(128 * jt);
            pass_order.push_back(o);
            cost_cum.push_back(cost_cum.back() + tile_cost_host[jt]);
        }
    }
    const int n_pass = (int)pass_feat0.size();

    // 3. batch of every 16-position chunk of a view, in the view's order (a chunk of padding only continues
    //    the batch before it, so that it never triggers a switch)
    const int n_chunks = n_pos / 16;
    std::vector<uint16_t> boc((size_t)V * n_chunks);
    for (int v = 0; v < V; ++v) {
        const int32_t* b = bos_host.data() + (size_t)v * M;
        const std::vector<int32_t>& pm = perms[view_order[v]];
        uint16_t last = 0;
        for (int c = 0; c < n_chunks; ++c) {
            if (pm[16 * c] >= 0) last = (uint16_t)b[pm[16 * c]];      // a chunk's first position is never padding unless all are
            boc[(size_t)v * n_chunks + c] = last;
        }
    }

    out.n_orders = n_orders; out.n_pass = n_pass; out.n_pos = n_pos;
    out.n_used = n_orders == 1 ? M : (int)n_used;
    out.perms = std::move(perms);
    out.view_order = std::move(view_order);
    out.pass_feat0 = std::move(pass_feat0);
    out.pass_order = std::move(pass_order);
    out.cost_cum = std::move(cost_cum);
    out.boc = std::move(boc);
}

int pmf_plan_batch_orders(int32_t M, int32_t N, int32_t n_views, const int32_t* cs, const int32_t* ce, const int32_t* nb,
                          const int32_t* bos, int32_t* n_orders, int32_t* n_pass, int32_t* n_pos, int32_t* view_order,
                          int32_t* perm, int64_t perm_cap, int32_t* pass_feat0, int32_t* pass_order, int32_t pass_cap,
                          uint16_t* chunk_batch, int64_t chunk_cap) {
    if (M <= 0 || N <= 0 || n_views <= 0 || !cs || !ce || !nb || !bos || !n_orders || !n_pass || !n_pos)
        return fail(nullptr, PMF_ERR_ARG, "bad arguments");
    std::vector<BatchView> views;
    int prev_end = 0;
    for (int v = 0; v < n_views; ++v) {
        if (cs[v] < prev_end || ce[v] > N || cs[v] >= ce[v] || nb[v] <= 0 || nb[v] > 65533)
            return fail(nullptr, PMF_ERR_ARG, "batch view %d: bad column range or batch count", v);
        prev_end = ce[v];
        views.push_back(BatchView{cs[v], ce[v], nb[v], 0});
        for (int i = 0; i < M; ++i)
            if (bos[(size_t)v * M + i] < 0 || bos[(size_t)v * M + i] >= nb[v])
                return fail(nullptr, PMF_ERR_ARG, "batch_of_sample[%d][%d] out of range", v, i);
    }
    std::vector<int32_t> b(bos, bos + (size_t)n_views * M), cost((N + 127) / 128, 1);
    TcBatchPlanHost pl;
    plan_tc_batches(M, round_up(M, 128), N, views, b, cost, pl);
    *n_orders = pl.n_orders; *n_pass = pl.n_pass; *n_pos = pl.n_pos;
    if (view_order) std::copy(pl.view_order.begin(), pl.view_order.end(), view_order);
    if (perm) {
        if (perm_cap < (int64_t)pl.n_orders * pl.n_pos) return fail(nullptr, PMF_ERR_ARG, "perm buffer too small");
        for (int o = 0; o < pl.n_orders; ++o) std::copy(pl.perms[o].begin(), pl.perms[o].end(), perm + (size_t)o * pl.n_pos);
    }
    if (pass_feat0 || pass_order) {
        if (pass_cap < pl.n_pass) return fail(nullptr, PMF_ERR_ARG, "pass buffers too small");
        if (pass_feat0) std::copy(pl.pass_feat0.begin(), pl.pass_feat0.end(), pass_feat0);
        if (pass_order) std::copy(pl.pass_order.begin(), pl.pass_order.end(), pass_order);
    }
    if (chunk_batch) {
        if (chunk_cap < (int64_t)pl.boc.size()) return fail(nullptr, PMF_ERR_ARG, "chunk buffer too small");
        std::copy(pl.boc.begin(), pl.boc.end(), chunk_batch);
    }
    return PMF_OK;
}

int pmf_model_s::build_tc_plan() {
    free_tc_plan();
    const int V = (int)views.size();
    const int n_jt = (N + 127) / 128;
    if (V == 0 || (int)tile_cost_host.size() != n_jt || bos_host.size() != (size_t)V * M)
        return fail(this, PMF_ERR_STATE, "batch layout / noise models are not set");
    for (const BatchView& bv : views)
        if (bv.n_batches > 65533) return fail(this, PMF_ERR_ARG, "tcgen05 data pass: more than 65533 batches in a view");

    TcBatchPlanHost pl;
    plan_tc_batches(M, Mp, N, views, bos_host, tile_cost_host, pl);
    const int n_orders = pl.n_orders, n_pass = pl.n_pass, n_pos = pl.n_pos;
    const size_t n_used = (size_t)pl.n_used;
    const std::vector<std::vector<int32_t>>& perms = pl.perms;
    const std::vector<int32_t>&view_order = pl.view_order, &pass_feat0 = pl.pass_feat0, &pass_order = pl.pass_order,
                              &cost_cum = pl.cost_cum;
    const std::vector<uint16_t>& boc = pl.boc;

    auto upload = [&](const void* src, size_t bytes, const void** dst) -> bool {
        void* d = nullptr;
        if (cudaMalloc(&d, bytes ? bytes : 4) != cudaSuccess) return false;
        tcb_allocs.push_back(d);
        if (src && bytes && copy_sync(stream, d, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return false;
        *dst = d;
        return true;
    };
    pmf::TcBatchDev& t = tcb;
    t.n_orders = n_orders; t.n_pass = n_pass; t.n_views = V;
    t.n_pos = n_pos; t.n_used = (int)n_used;
    t.direct = n_orders == 1;
    bool ok = upload(pass_feat0.data(), (size_t)n_pass * 4, (const void**)&t.pass_feat0) &&
              upload(pass_order.data(), (size_t)n_pass * 4, (const void**)&t.pass_order) &&
              upload(view_order.data(), (size_t)V * 4, (const void**)&t.view_order) &&
              upload(cost_cum.data(), ((size_t)n_pass + 1) * 4, (const void**)&t.cost_cum) &&
              upload(boc.data(), boc.size() * 2, (const void**)&t.boc);
    if (ok && !t.direct) {
        std::vector<int32_t> perm_flat((size_t)n_orders * n_pos), pos_flat((size_t)n_orders * M);
        for (int o = 0; o < n_orders; ++o)
            for (int p = 0; p < n_pos; ++p) {
                perm_flat[(size_t)o * n_pos + p] = perms[o][p];
                if (perms[o][p] >= 0) pos_flat[(size_t)o * M + perms[o][p]] = p;
            }
        const size_t rows = (size_t)n_orders * n_pos;
        ok = upload(perm_flat.data(), perm_flat.size() * 4, (const void**)&t.perm) &&
             upload(pos_flat.data(), pos_flat.size() * 4, (const void**)&t.pos) &&
             upload(nullptr, rows * 64 * 4, (const void**)&t.Xh) && upload(nullptr, rows * 64 * 4, (const void**)&t.Xb) &&
             upload(nullptr, rows * Kp * 4, (const void**)&t.dX);
        if (ok) {
            float* a = nullptr;
            ok = big_alloc(&a, (size_t)n_pass * 128 * n_pos, dims.device) == cudaSuccess;
            t.A_tc = a;
        }
        if (ok) {
            // operand scratch is 64 / 128 wide and zero beyond Kp; the dX copies start (and are left) clean
            ok = cudaMemsetAsync(t.Xh, 0, rows * 64 * 4, stream) == cudaSuccess &&
                 cudaMemsetAsync(t.Xb, 0, rows * 64 * 4, stream) == cudaSuccess &&
                 cudaMemsetAsync(t.dX, 0, rows * Kp * 4, stream) == cudaSuccess;
            pmf::DataPassParams p;
            std::memset(&p, 0, sizeof p);
            p.N = N; p.lda = lda; p.Mp = Mp; p.A = A; p.bcol_view = bcol_view;
            ok = ok && pmf::launch_build_a_tc(p, t, const_cast<float*>(t.A_tc), stream) == cudaSuccess;
            ++launches;
        }
    }
    if (!ok) {
        free_tc_plan();
        return fail(this, PMF_ERR_ALLOC, "device allocation failed (tcgen05 batch plan: %d orders, %d passes)", n_orders, n_pass);
    }
    tcb_valid = true;
    return 0;
}

int pmf_model_s::run_data_pass(DataPassParams& p, int kind, int precision) {
    bool batches_ok = true;                    // chunk tables hold 16-bit batch ids
    for (const BatchView& bv : views) batches_ok &= bv.n_batches <= 65533;
    const bool tc_ok = tc_supported(p) && cc_major == 10 && batches_ok;
    const bool wide_ok = wide_supported(p) && cc_major == 10;
    if (kind == PMF_KERNEL_TC && !tc_ok && !wide_ok)
        return fail(this, PMF_ERR_ARG, "tcgen05 data pass needs an sm_100 device and K <= 64 (at most 65533 batches per "
                                       "view) or 64 < K <= 256 without batch layers");
    // AUTO: the tensor-core kernels take a problem when it is large enough for their single-pass TF32 gradient
    // contractions to stay below the 1e-4 parity bar; smaller ones run the exact-FP32 kernel.  Measured against the
    // FP32 kernel (scripts/tc_precision_vs_size.py, profiles/r2_tc_precision_vs_size.jsonl; freshly initialised
    // model): K = 64: 8.8e-5 at 2 000 x 3 000, 6e-5 at 4 000 x 6 000, 3-4e-5 at 10 000 x 30 000; K = 128: 1.2e-4 at
    // 2 000 x 3 000, 8.8e-5 at 4 000 x 6 000; K = 256: 1.2e-4 at 4 000 x 6 000, 6-8e-5 at 10 000 x 30 000 -- the error
    // goes with K / sqrt(M N), hence the K^2 in the size rule.  (Near a fitted model the gradients shrink and the
    // rounding noise does not: 5-7e-5 at the C2 shape after 30 epochs.)
    const double kscale = std::max(1.0, Kp / 64.0);
    const bool big = (double)M * (double)N >= 6.0e6 * kscale * kscale && std::min(M, N) >= 2000;
    if (wide_ok && (kind == PMF_KERNEL_TC || (kind == PMF_KERNEL_AUTO && auto_tc && big))) {
        if (!wide.G) {
            const size_t nx = wide_scratch_floats(Mp, Kp), ny = wide_scratch_floats(Np, Kp);
            float *xb = nullptr, *yb = nullptr;
            if (big_alloc(&wide.G, (size_t)N * lda, dims.device) != cudaSuccess || dev_alloc(&wide.Xh, nx) != cudaSuccess ||
                dev_alloc(&xb, nx) != cudaSuccess || dev_alloc(&wide.Yh, ny) != cudaSuccess || dev_alloc(&yb, ny) != cudaSuccess)
                return fail(this, PMF_ERR_ALLOC, "device allocation failed (K > 64 tensor-core path needs a second %.2f GB matrix)",
                            (double)N * lda * 4e-9);
            wide.Xb = xb; wide.Yb = yb;
            // G' rows are written for every column j < N and every sample position i < lda; nothing else is ever read
        }
        cudaEvent_t t0 = nullptr, t1 = nullptr;
        if (profiling) {
            if (prof_used + 2 > prof_ev.size()) {
                for (int i = 0; i < 2; ++i) { cudaEvent_t ev; cudaEventCreate(&ev); prof_ev.push_back(ev); }
            }
            t0 = prof_ev[prof_used]; t1 = prof_ev[prof_used + 1];
            prof_used += 2;
            cudaEventRecord(t0, stream);
        }
        int n_launched = 0;
        cudaError_t e = launch_data_pass_wide(p, wide, precision, stream, n_sms, &n_launched);
        if (profiling) cudaEventRecord(t1, stream);
        if (e != cudaSuccess) { cuda_failed = true; return fail(this, PMF_ERR_CUDA, "tcgen05 (K > 64) data pass launch: %s", cudaGetErrorString(e)); }
        launches += n_launched;
        xsplit_valid = false;
        return 0;
    }
    const bool use_tc = kind == PMF_KERNEL_TC || (kind == PMF_KERNEL_AUTO && tc_ok && auto_tc && big);
    if (use_tc) {
        if (!views.empty() && !tcb_valid) {
            int rc = build_tc_plan();
            if (rc != 0) return rc;
        }
        if (!Xh) {
            // 64-wide FP32 and 128-wide BF16 operand scratch, zero padded beyond Kp (never written there)
            if (dev_alloc(&Xh, (size_t)Mp * 64) != cudaSuccess || dev_alloc(&Xl, (size_t)Mp * 64) != cudaSuccess ||
                cudaMemsetAsync(Xh, 0, (size_t)Mp * 64 * 4, stream) != cudaSuccess ||
                cudaMemsetAsync(Xl, 0, (size_t)Mp * 64 * 4, stream) != cudaSuccess)
                return fail(this, PMF_ERR_ALLOC, "device allocation failed");
        }
        cudaEvent_t t0 = nullptr, t1 = nullptr;
        if (profiling) {
            if (prof_used + 2 > prof_ev.size()) {
                for (int i = 0; i < 2; ++i) { cudaEvent_t ev; cudaEventCreate(&ev); prof_ev.push_back(ev); }
            }
            t0 = prof_ev[prof_used]; t1 = prof_ev[prof_used + 1];
            prof_used += 2;
            cudaEventRecord(t0, stream);
        }
        const bool refresh = !xsplit_valid;
        int n_launched = 0;
        cudaError_t e = launch_data_pass_tc(p, Xh, Xl, refresh, precision, stream, n_sms, views.empty() ? nullptr : &tcb, &n_launched);
        xsplit_valid = true;     // stays true only while every later change of X goes through run_update_multi
        if (profiling) cudaEventRecord(t1, stream);
        if (e != cudaSuccess) { cuda_failed = true; return fail(this, PMF_ERR_CUDA, "tcgen05 data pass launch: %s", cudaGetErrorString(e)); }
        launches += n_launched;
        return 0;
    }
    xsplit_valid = false;
    size_t smem = sizeof(float) * ((size_t)128 * (Kp + 4) + 64 * 68 + 2 * (size_t)64 * nb_max);
    if (smem > 227 * 1024) return fail(this, PMF_ERR_ARG, "too many batches per view (%d) for K=%d", nb_max, K);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (profiling) {
        if (prof_used + 2 > prof_ev.size()) {
            for (int i = 0; i < 2; ++i) { cudaEvent_t ev; cudaEventCreate(&ev); prof_ev.push_back(ev); }
        }
        e0 = prof_ev[prof_used]; e1 = prof_ev[prof_used + 1];
        prof_used += 2;
        cudaEventRecord(e0, stream);
    }
    cudaError_t e = launch_data_pass_ffma(p, stream, n_sms);
    if (profiling) cudaEventRecord(e1, stream);
    if (e != cudaSuccess) { cuda_failed = true; return fail(this, PMF_ERR_CUDA, "data pass launch: %s", cudaGetErrorString(e)); }
    launches++;
    return 0;
}

void pmf_model_s::fill_factor_params(int which, FactorUpdateParams& q) {
    std::memset(&q, 0, sizeof q);
    q.n = which == 0 ? M : N; q.Kp = Kp; q.K = K;
    q.P = which == 0 ? X : Y;
    q.grad = which == 0 ? dX : g_Y();
    q.acc = which == 0 ? accX : accY;
}

int pmf_model_s::run_network_reg(int which, const int* stop) {
    SideReg& r = reg[which];
    if (!r.net.present) return 0;
    NetworkParams q;
    std::memset(&q, 0, sizeof q);
    q.n = which == 0 ? M : N; q.Kp = Kp; q.K = K;
    q.P = which == 0 ? X : Y; q.grad = which == 0 ? dX : g_Y();
    auto cv = [](const DevCsr& d) { CsrBlock b{d.rowptr, d.col, d.val, d.rowptr_base, d.nnz_base}; return b; };
    q.AA = cv(r.net.AA); q.AB = cv(r.net.AB); q.BB = cv(r.net.BB); q.ABt = cv(r.net.ABt);
    q.nv = r.net.nv; q.virt_base = r.net.virt_base; q.u = r.net.u; q.work = r.net.work; q.nv_total = r.net.nv_total;
    q.yt = r.net.yt; q.gt = r.net.gt; q.ld = r.net.ld;
    q.p = r.net.p; q.rtol = r.net.rtol; q.atol = r.net.atol; q.itmax = r.net.itmax;
    q.loss_out = scalars + (which == 0 ? SC_XREG : SC_YREG); q.stop_flag = stop;
    cudaError_t e = launch_network_reg(q, stream, n_sms, r.net.nv_max);
    if (e != cudaSuccess) { cuda_failed = true; return fail(this, PMF_ERR_CUDA, "network reg launch: %s", cudaGetErrorString(e)); }
    launches += r.net.nv_max > 0 ? 4 : 3;
    return 0;
}

static void fill_factor_reg(pmf_model_s* h, int which, FactorUpdateParams& q, const int* stop) {
    SideReg& r = h->reg[which];
    h->fill_factor_params(which, q);
    q.grad_out = which == 0 ? h->dX : h->g_Y();      // in-place: data gradient + penalty pullback
    q.l2_w = r.l2_w; q.group_id = r.group_id; q.group_w = r.group_w;
    q.l1_mask = r.l1_mask; q.l1_w = r.l1_w;
    q.ard_alpha = r.ard_alpha; q.ard_beta_row = r.ard_beta_row; q.ard_beta_full = r.ard_beta_full;
    q.loss_out = h->scalars + (which == 0 ? SC_XREG : SC_YREG); q.do_update = 0; q.stop_flag = stop;
}

// segments of the vector-parameter block: slot 1 logsigma, slot 3 mu, slot 2 logdelta, slot 4 theta
struct VecSeg { size_t off; int n; int slot; };
static void vec_segments(pmf_model_s* h, VecSeg segs[4]) {
    segs[0] = {0, h->N, 1};
    segs[1] = {(size_t)h->Np, h->N, 3};
    segs[2] = {2 * (size_t)h->Np, (int)h->nbp, 2};
    segs[3] = {2 * (size_t)h->Np + (size_t)h->nbp, (int)h->nbp, 4};
}

// every elementwise penalty of the requested sides in ONE launch
int pmf_model_s::run_reg_multi(bool x_side, bool y_side, bool vectors, const int* stop) {
    MultiPassParams mp;
    std::memset(&mp, 0, sizeof mp);
    if (x_side && reg[0].any_elementwise()) fill_factor_reg(this, 0, mp.f[mp.nf++], stop);
    if (y_side && reg[1].any_elementwise()) fill_factor_reg(this, 1, mp.f[mp.nf++], stop);
    if (vectors) {
        VecSeg segs[4];
        vec_segments(this, segs);
        for (const VecSeg& sgm : segs) {
            const unsigned bit = 1u << (sgm.slot - 1);
            // a frozen layer's ColParamReg / BatchArrayReg evaluates to 0 in the reference (src/regularizers.jl:509, :867)
            if (sgm.n <= 0 || !layer_reg_present[sgm.slot - 1] || ((frozen_regs | frozen_layers) & bit)) continue;
            VectorUpdateParams& q = mp.v[mp.nv++];
            q.n = sgm.n; q.p = vp + sgm.off; q.grad = sg + (size_t)Np * Kp + sgm.off; q.acc = accvp + sgm.off;
            q.stop_flag = stop;
            q.reg_w = regw + sgm.off; q.reg_c = regc + sgm.off; q.reg_active = 1;
            q.grad_out = sg + (size_t)Np * Kp + sgm.off;
            q.loss_out = scalars + SC_LAYERREG;
        }
    }
    if (mp.nf == 0 && mp.nv == 0) return 0;
    cudaError_t e = launch_multi_pass(mp, stream, n_sms);
    if (e != cudaSuccess) { cuda_failed = true; return fail(this, PMF_ERR_CUDA, "penalty pass launch: %s", cudaGetErrorString(e)); }
    launches++;
    return 0;
}

// AdaGrad step of every enabled parameter array in ONE launch.  The same pass clears the gradient
// buffers and the loss scalars for the next epoch and, for X, writes the TF32 operand split of the
// updated values, so the next epoch starts directly with the data pass.
int pmf_model_s::run_update_multi(bool upd_X, bool upd_Y, bool upd_layers, bool upd_noise, float lr, float eps, const int* stop) {
    MultiPassParams mp;
    std::memset(&mp, 0, sizeof mp);
    {
        FactorUpdateParams& q = mp.f[mp.nf++];
        fill_factor_params(0, q);
        q.do_update = upd_X ? 1 : 0; q.lr = lr; q.eps = eps; q.stop_flag = stop;
        q.zero_buf = dX; q.n_pad = Mp;
        if (Xh && xsplit_valid) { q.Ph = Xh; q.Pl = Xl; }
    }
    {
        FactorUpdateParams& q = mp.f[mp.nf++];
        fill_factor_params(1, q);
        q.do_update = upd_Y ? 1 : 0; q.lr = lr; q.eps = eps; q.stop_flag = stop;
        q.zero_buf = g_Y(); q.n_pad = Np;
    }
    VecSeg segs[4];
    vec_segments(this, segs);
    for (const VecSeg& sgm : segs) {
        if (sgm.n <= 0) continue;
        const unsigned bit = 1u << (sgm.slot - 1);
        VectorUpdateParams& q = mp.v[mp.nv++];
        q.n = sgm.n; q.p = vp + sgm.off; q.grad = sg + (size_t)Np * Kp + sgm.off; q.acc = accvp + sgm.off;
        q.stop_flag = stop;
        q.do_update = (upd_layers && !(frozen_layers & bit)) ? 1 : 0; q.lr = lr; q.eps = eps;
        q.zero_buf = sg + (size_t)Np * Kp + sgm.off;
    }
    mp.zero_scalars = scalars;
    if (has_ordinal) {
        mp.thr = thresholds; mp.thr_grad = g_thr(); mp.thr_acc = acc_thr; mp.thr_ranges = n_ranges;
        mp.thr_update = upd_noise ? 1 : 0; mp.thr_lr = lr; mp.thr_eps = eps;
    }
    cudaError_t e = launch_multi_pass(mp, stream, n_sms);
    if (e != cudaSuccess) { cuda_failed = true; return fail(this, PMF_ERR_CUDA, "update pass launch: %s", cudaGetErrorString(e)); }
    launches++;
    grads_clean = true;     // valid for the next epoch of this fit only (pmf_fit_start / pmf_loss_grad reset it)
    return 0;
}


// ---- fused epoch pass (pmf_fit) --------------------------------------------------------------------------------
// Every elementwise penalty and every AdaGrad step of the epoch as ONE MultiPassParams: the segments of run_reg_multi
// and run_update_multi merged.  Penalty values are taken at the UPDATED parameters and added into `scalars` (which the
// caller has pointed at the next epoch's buffer).  values_only: no gradients touched, no update, values at the
// current parameters (the penalties of a fit's first epoch).
void pmf_model_s::fill_epoch_pass(MultiPassParams& mp, const pmf_fit_opts* o, bool values_only) {
    std::memset(&mp, 0, sizeof mp);
    const bool upd[2] = {o->update_X != 0, o->update_Y != 0};
    for (int which = 0; which < 2; ++which) {
        FactorUpdateParams& q = mp.f[mp.nf++];
        fill_factor_params(which, q);
        SideReg& r = reg[which];
        q.l2_w = r.l2_w; q.group_id = r.group_id; q.group_w = r.group_w;
        q.l1_mask = r.l1_mask; q.l1_w = r.l1_w;
        q.ard_alpha = r.ard_alpha; q.ard_beta_row = r.ard_beta_row; q.ard_beta_full = r.ard_beta_full;
        q.loss_out = r.any_elementwise() ? scalars + (which == 0 ? SC_XREG : SC_YREG) : nullptr;
        q.loss_after = 1;
        q.lr = o->lr; q.eps = o->adagrad_eps;
        if (!values_only) {
            q.do_update = upd[which] ? 1 : 0;
            q.zero_buf = which == 0 ? dX : g_Y(); q.n_pad = which == 0 ? Mp : Np;
            if (which == 0 && Xh && xsplit_valid) { q.Ph = Xh; q.Pl = Xl; }
        }
    }
    VecSeg segs[4];
    vec_segments(this, segs);
    for (const VecSeg& sgm : segs) {
        if (sgm.n <= 0) continue;
        const unsigned bit = 1u << (sgm.slot - 1);
        // a frozen layer's (or frozen regulariser's) penalty evaluates to 0 (src/regularizers.jl:508-510, :887, :950-1004)
        const bool reg_on = layer_reg_present[sgm.slot - 1] && !((frozen_regs | frozen_layers) & bit);
        if (values_only && !reg_on) continue;
        VectorUpdateParams& q = mp.v[mp.nv++];
        q.n = sgm.n; q.p = vp + sgm.off; q.grad = sg + (size_t)Np * Kp + sgm.off; q.acc = accvp + sgm.off;
        q.reg_w = regw + sgm.off; q.reg_c = regc + sgm.off; q.reg_active = reg_on ? 1 : 0;
        q.loss_out = reg_on ? scalars + SC_LAYERREG : nullptr;
        q.loss_after = 1;
        q.lr = o->lr; q.eps = o->adagrad_eps;
        if (!values_only) {
            q.do_update = (o->update_col_layers && !(frozen_layers & bit)) ? 1 : 0;
            q.zero_buf = sg + (size_t)Np * Kp + sgm.off;
        }
    }
    if (!values_only && has_ordinal) {
        mp.thr = thresholds; mp.thr_grad = g_thr(); mp.thr_acc = acc_thr; mp.thr_ranges = n_ranges;
        mp.thr_update = o->update_noise_models ? 1 : 0; mp.thr_lr = o->lr; mp.thr_eps = o->adagrad_eps;
    }
}

int pmf_model_s::run_penalty_values(const pmf_fit_opts* o) {
    MultiPassParams mp;
    fill_epoch_pass(mp, o, true);
    // only segments that carry a penalty
    MultiPassParams pr;
    std::memset(&pr, 0, sizeof pr);
    for (int i = 0; i < mp.nf; ++i)
        if (mp.f[i].loss_out) pr.f[pr.nf++] = mp.f[i];
    for (int i = 0; i < mp.nv; ++i) pr.v[pr.nv++] = mp.v[i];
    if (pr.nf == 0 && pr.nv == 0) return 0;
    cudaError_t e = launch_multi_pass(pr, stream, n_sms);
    if (e != cudaSuccess) { cuda_failed = true; return fail(this, PMF_ERR_CUDA, "penalty value pass launch: %s", cudaGetErrorString(e)); }
    launches++;
    return 0;
}

int pmf_model_s::run_fused_epoch(const pmf_fit_opts* o, const FusedControl& fc) {
    MultiPassParams mp;
    fill_epoch_pass(mp, o, false);
    cudaError_t e = launch_fused_epoch_pass(mp, fc, stream, n_sms);
    if (e != cudaSuccess) { cuda_failed = true; return fail(this, PMF_ERR_CUDA, "epoch pass launch: %s", cudaGetErrorString(e)); }
    launches++;
    grads_clean = true;
    return 0;
}
