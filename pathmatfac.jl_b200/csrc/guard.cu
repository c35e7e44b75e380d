// The library's device allocator: every cudaMalloc / cudaFree of libpmf lands here (pmf_internal.h).
//
// (1) A size-keyed cache of freed blocks.  mf_fit! on a host-resident model creates and destroys a handle per call
// (src/fit.jl:9-38 with gpu() / cpu() around it): ~30 cudaMalloc + ~30 cudaFree of 4 B - 8 MB each, measured at 5-6 ms
// + 3-5 ms of a 43 ms call at the C2 shape (profiles/r2_e2e_breakdown.log), with occasional stalls of hundreds of
// milliseconds inside the driver.  A freed block of at most 64 MB is parked (up to 1 GB / 4096 blocks per process)
// and handed to the next allocation of exactly the same size on the same device, zero-filled -- the state a fresh
// cudaMalloc block has in practice -- after a device synchronisation at the free, which is what cudaFree implies.
// Larger blocks (the data matrix and its copies) have their own two-slot pool in pmf_abi.cu.
// pmf_release_cached_memory() returns everything to the driver; PMF_ALLOC_CACHE=0 disables the cache.
//
// (2) Guard zones around every device allocation (test hook, enabled by PMF_GUARD=1 in the environment before the
// first allocation; the cache is off then): 1 KiB of a byte pattern in front of and behind each buffer, verified by pmf_check_guards().  The
// pool's compute-sanitizer is closed, so out-of-bounds writes of the plain-pointer code (atomic flushes, operand
// splits, gathers; TMA accesses are bounds-checked by their tensor maps) are caught this way in the GPU tests.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <map>
#include <mutex>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/pmf.h"

namespace pmf {

namespace {
constexpr size_t kGuard = 1024;
constexpr unsigned char kPattern = 0xA5;
struct Rec { void* base; size_t bytes; int dev; };
std::mutex g_mu;
std::map<void*, Rec> g_allocs;     // user pointer -> allocation
bool enabled() {
    static const bool on = [] { const char* e = getenv("PMF_GUARD"); return e && e[0] == '1'; }();
    return on;
}

// ---- cache of freed blocks -----------------------------------------------------------------------
constexpr size_t kCacheBlockMax = 64u << 20;
constexpr size_t kCacheTotalMax = 1u << 30;
constexpr size_t kCacheCountMax = 4096;
std::mutex g_cache_mu;
std::multimap<std::pair<int, size_t>, void*> g_cache;               // (device, bytes) -> parked block
std::unordered_map<void*, std::pair<int, size_t>> g_live;           // live cacheable block -> (device, bytes)
size_t g_cache_bytes = 0;
bool cache_enabled() {
    static const bool on = [] { const char* e = getenv("PMF_ALLOC_CACHE"); return !(e && e[0] == '0'); }();
    return on && !enabled();
}
}  // namespace

void release_alloc_cache() {
    std::vector<std::pair<int, void*>> drop;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        for (const auto& kv : g_cache) drop.emplace_back(kv.first.first, kv.second);
        g_cache.clear();
        g_cache_bytes = 0;
    }
    if (drop.empty()) return;
    int cur = 0;
    cudaGetDevice(&cur);
    for (const auto& d : drop) {
        cudaSetDevice(d.first);
        cudaFree(d.second);
    }
    cudaSetDevice(cur);
}

static cudaError_t cached_malloc(void** p, size_t bytes) {
    *p = nullptr;
    const bool cacheable = bytes > 0 && bytes <= kCacheBlockMax;
    int dev = 0;
    if (cacheable) {
        cudaGetDevice(&dev);
        void* hit = nullptr;
        {
            std::lock_guard<std::mutex> lk(g_cache_mu);
            auto it = g_cache.find(std::make_pair(dev, bytes));
            if (it != g_cache.end()) {
                hit = it->second;
                g_cache.erase(it);
                g_cache_bytes -= bytes;
                g_live[hit] = std::make_pair(dev, bytes);
            }
        }
        if (hit) {
            // zero-filled and complete before the caller sees the block (the handle's stream is non-blocking, so a
            // pending legacy-stream memset would not be ordered against its work)
            cudaError_t e = cudaMemsetAsync(hit, 0, bytes, 0);
            if (e == cudaSuccess) e = cudaStreamSynchronize(0);
            if (e != cudaSuccess) return e;
            *p = hit;
            return cudaSuccess;
        }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaErrorMemoryAllocation) {      // give the parked blocks back and retry once
        cudaGetLastError();
        release_alloc_cache();
        e = cudaMalloc(p, bytes);
    }
    if (e == cudaSuccess && cacheable) {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        g_live[*p] = std::make_pair(dev, bytes);
    }
    return e;
}

static cudaError_t cached_free(void* p) {
    if (p == nullptr) return cudaSuccess;
    std::pair<int, size_t> rec(0, 0);
    bool known = false;
    {
        std::lock_guard<std::mutex> lk(g_cache_mu);
        auto it = g_live.find(p);
        if (it != g_live.end()) {
            rec = it->second;
            g_live.erase(it);
            known = g_cache_bytes + rec.second <= kCacheTotalMax && g_cache.size() < kCacheCountMax;
        }
    }
    if (!known) return cudaFree(p);
    // cudaFree implies a device synchronisation: nothing may still use the block when another allocation takes it
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFree(p); return e; }
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_cache.emplace(rec, p);
    g_cache_bytes += rec.second;
    return cudaSuccess;
}

cudaError_t guarded_malloc(void** p, size_t bytes) {
    if (!enabled()) return cache_enabled() ? cached_malloc(p, bytes) : cudaMalloc(p, bytes);
    void* base = nullptr;
    cudaError_t e = cudaMalloc(&base, bytes + 2 * kGuard);
    if (e != cudaSuccess) { *p = nullptr; return e; }
    cudaMemset(base, kPattern, kGuard);
    cudaMemset(static_cast<char*>(base) + kGuard + bytes, kPattern, kGuard);
    cudaStreamSynchronize(0);      // the pattern is in place before any (non-blocking) stream touches the block
    *p = static_cast<char*>(base) + kGuard;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_mu);
    g_allocs[*p] = Rec{base, bytes, dev};
    return cudaSuccess;
}

cudaError_t guarded_free(void* p) {
    if (!enabled()) return cache_enabled() ? cached_free(p) : cudaFree(p);
    if (p == nullptr) return cudaSuccess;
    void* base = p;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_allocs.find(p);
        if (it != g_allocs.end()) { base = it->second.base; g_allocs.erase(it); }
    }
    return cudaFree(base);
}

}  // namespace pmf

extern "C" int pmf_check_guards(int64_t* n_buffers, int64_t* n_corrupt_bytes) {
    int64_t nb = 0, bad = 0;
    if (pmf::enabled()) {
        cudaDeviceSynchronize();
        std::vector<unsigned char> host(2 * pmf::kGuard);
        std::lock_guard<std::mutex> lk(pmf::g_mu);
        int cur = 0;
        cudaGetDevice(&cur);
        for (const auto& kv : pmf::g_allocs) {
            const pmf::Rec& r = kv.second;
            cudaSetDevice(r.dev);
            cudaMemcpy(host.data(), r.base, pmf::kGuard, cudaMemcpyDeviceToHost);
            cudaMemcpy(host.data() + pmf::kGuard, static_cast<char*>(r.base) + pmf::kGuard + r.bytes, pmf::kGuard, cudaMemcpyDeviceToHost);
            for (unsigned char c : host) bad += c != pmf::kPattern;
            ++nb;
        }
        cudaSetDevice(cur);
    }
    if (n_buffers) *n_buffers = nb;
    if (n_corrupt_bytes) *n_corrupt_bytes = bad;
    return pmf::enabled() ? PMF_OK : PMF_ERR_STATE;
}
