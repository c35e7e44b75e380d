// Drives pmf::guarded_malloc / guarded_free / release_alloc_cache (csrc/guard.cu, compiled against the stub runtime
// in this directory) and prints one "name value" line per observation; tests/test_allocator_cpu.py asserts on them.
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include "cuda_runtime.h"

StubCounters g_stub = {0, 0, 0, 0, 0, 0, 0, 0};

namespace pmf {
cudaError_t guarded_malloc(void** p, size_t bytes);
cudaError_t guarded_free(void* p);
void release_alloc_cache();
}
extern "C" int pmf_check_guards(int64_t* n_buffers, int64_t* n_corrupt_bytes);

static bool all_zero(const void* p, size_t n) {
    const unsigned char* c = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; ++i) if (c[i]) return false;
    return true;
}

int main(int argc, char** argv) {
    const bool guard_mode = argc > 1 && argv[1][0] == 'g';
    void *a = nullptr, *b = nullptr, *c = nullptr;
    pmf::guarded_malloc(&a, 1000);
    pmf::guarded_malloc(&b, 1000);
    pmf::guarded_malloc(&c, 4096);
    printf("mallocs_after_three %ld\n", g_stub.mallocs);
    if (guard_mode) {
        int64_t nb = 0, bad = 0;
        int rc = pmf_check_guards(&nb, &bad);
        printf("guard_rc %d\nguard_buffers %lld\nguard_bad %lld\n", rc, (long long)nb, (long long)bad);
        static_cast<unsigned char*>(a)[1000] = 0;          // one byte past the end
        static_cast<unsigned char*>(c)[-1] = 0;            // one byte before the start
        pmf_check_guards(&nb, &bad);
        printf("guard_bad_after_overrun %lld\n", (long long)bad);
        pmf::guarded_free(a); pmf::guarded_free(b); pmf::guarded_free(c);
        printf("frees_in_guard_mode %ld\n", g_stub.frees);
        return 0;
    }
    memset(a, 0x77, 1000);
    pmf::guarded_free(a);
    printf("frees_after_cached_free %ld\ndevice_syncs_after_cached_free %ld\n", g_stub.frees, g_stub.device_syncs);
    void* a2 = nullptr;
    pmf::guarded_malloc(&a2, 1000);                        // same size, same device: the parked block, zero-filled
    printf("reused_same_block %d\nreused_block_is_zero %d\nmallocs_after_reuse %ld\n", a2 == a, all_zero(a2, 1000), g_stub.mallocs);
    pmf::guarded_free(a2);
    void* d = nullptr;
    pmf::guarded_malloc(&d, 1001);                         // another size: a fresh block
    printf("other_size_is_fresh %d\n", d != a && g_stub.mallocs == 4);
    g_stub.device = 1;
    void* e = nullptr;
    pmf::guarded_malloc(&e, 1000);                         // same size on ANOTHER device: not the parked block
    printf("other_device_is_fresh %d\n", e != a && g_stub.mallocs == 5);
    g_stub.device = 0;
    // blocks above 64 MB are never parked
    void* big = nullptr;
    const size_t big_n = (64u << 20) + 1;
    pmf::guarded_malloc(&big, big_n);
    long f0 = g_stub.frees;
    pmf::guarded_free(big);
    printf("big_block_freed_at_once %d\n", g_stub.frees == f0 + 1);
    // the cache holds at most 1 GB: twenty 60 MB blocks -> seventeen parked (with the small ones), the rest freed
    std::vector<void*> blocks(20);
    for (auto& q : blocks) pmf::guarded_malloc(&q, 60u << 20);
    f0 = g_stub.frees;
    for (auto& q : blocks) pmf::guarded_free(q);
    printf("frees_beyond_the_cap %ld\n", g_stub.frees - f0);
    // an allocation failure gives the parked blocks back and retries
    g_stub.fail_next_mallocs = 1;
    void* r = nullptr;
    f0 = g_stub.frees;
    cudaError_t er = pmf::guarded_malloc(&r, 12345);
    printf("retry_after_oom_ok %d\nparked_blocks_released_on_oom %d\n", er == cudaSuccess && r != nullptr, g_stub.frees - f0 >= 17);
    pmf::guarded_free(r); pmf::guarded_free(b); pmf::guarded_free(c); pmf::guarded_free(d);
    g_stub.device = 1; pmf::guarded_free(e); g_stub.device = 0;
    pmf::release_alloc_cache();
    printf("live_bytes_after_release %zu\nmallocs_equal_frees %d\n", g_stub.live_bytes, g_stub.mallocs == g_stub.frees);
    pmf::guarded_free(nullptr);
    printf("free_null_ok 1\n");
    return 0;
}
