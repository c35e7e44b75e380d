// Guard zones around every device allocation of libpmf (test hook, enabled by PMF_GUARD=1 in the environment before the
// first allocation): 1 KiB of a byte pattern in front of and behind each buffer, verified by pmf_check_guards().  The
// pool's compute-sanitizer is closed, so out-of-bounds writes of the plain-pointer code (atomic flushes, operand
// splits, gathers; TMA accesses are bounds-checked by their tensor maps) are caught this way in the GPU tests.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <map>
#include <mutex>
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
}  // namespace

cudaError_t guarded_malloc(void** p, size_t bytes) {
    if (!enabled()) return cudaMalloc(p, bytes);
    void* base = nullptr;
    cudaError_t e = cudaMalloc(&base, bytes + 2 * kGuard);
    if (e != cudaSuccess) { *p = nullptr; return e; }
    cudaMemset(base, kPattern, kGuard);
    cudaMemset(static_cast<char*>(base) + kGuard + bytes, kPattern, kGuard);
    *p = static_cast<char*>(base) + kGuard;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_mu);
    g_allocs[*p] = Rec{base, bytes, dev};
    return cudaSuccess;
}

cudaError_t guarded_free(void* p) {
    if (!enabled() || p == nullptr) return cudaFree(p);
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
