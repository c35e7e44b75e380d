// Host-only stand-in for the handful of CUDA runtime calls csrc/guard.cu makes, so that the allocator's bookkeeping
// (size-keyed cache, caps, release, guard zones) can be unit-tested without a GPU (tests/test_allocator_cpu.py).
// "Device" memory is host memory; every call is counted.
#pragma once
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

typedef enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 } cudaError_t;
typedef enum { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2 } cudaMemcpyKind;
typedef void* cudaStream_t;

struct StubCounters { long mallocs, frees, memsets, syncs, device_syncs; size_t live_bytes; long fail_next_mallocs; int device; };
extern StubCounters g_stub;

static inline cudaError_t cudaMalloc(void** p, size_t n) {
    if (g_stub.fail_next_mallocs > 0) { --g_stub.fail_next_mallocs; *p = nullptr; return cudaErrorMemoryAllocation; }
    size_t* q = static_cast<size_t*>(malloc(n + 16));
    q[0] = n;
    *p = reinterpret_cast<char*>(q) + 16;
    memset(*p, 0xCD, n);            // fresh "device" memory is NOT zero here: the cache must not rely on it
    ++g_stub.mallocs; g_stub.live_bytes += n;
    return cudaSuccess;
}
static inline cudaError_t cudaFree(void* p) {
    if (!p) return cudaSuccess;
    size_t* q = reinterpret_cast<size_t*>(static_cast<char*>(p) - 16);
    ++g_stub.frees; g_stub.live_bytes -= q[0];
    free(q);
    return cudaSuccess;
}
static inline cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); ++g_stub.memsets; return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { return cudaMemset(p, v, n); }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { ++g_stub.syncs; return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { ++g_stub.device_syncs; return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = g_stub.device; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int d) { g_stub.device = d; return cudaSuccess; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
