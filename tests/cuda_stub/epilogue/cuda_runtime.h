// Host-only stand-in that lets g++ compile csrc/pmf_epilogue.cuh -- the per-entry noise-model math the FP32 and the
// tensor-core kernels share -- so that tests/test_epilogue_cpu.py can evaluate the SHIPPED source against the oracle
// without a GPU.  The fast-math intrinsics map to libm (on the device they are single MUFU instructions with ~2 ulp
// of error; the GPU suite measures that against the same oracle), the warp / atomic primitives of the reduction helpers
// are declared but never called from the harness.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __expf(x) expf(x)       // (glibc declares these two names itself: macros, not functions)
#define __logf(x) logf(x)
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline unsigned __activemask() { return 1u; }
static inline int __all_sync(unsigned, int p) { return p; }
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
static inline float atomicAdd(float* p, float v) { float o = *p; *p += v; return o; }
static inline void __syncthreads() {}
struct Dim3Stub { unsigned x, y, z; };
static const Dim3Stub threadIdx = {0, 0, 0}, blockDim = {1, 1, 1};
