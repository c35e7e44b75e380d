// Micro-benchmark: how fast can 148 persistent CTAs stream the data matrix A through TMA in the access pattern of the
// fused data pass (128-feature x 64-sample tiles of a [N][lda] FP32 array = 2 boxes of 128 rows x 128 bytes), as a
// function of the ring depth, against contiguous 32 KB bulk copies of a tile-blocked copy of the same bytes?
// Nothing is computed: one thread per CTA issues the loads, one thread per stage waits for its tile, holds the stage
// for `hold` cycles (the epilogue + MMA2 of the data pass) and hands it back: period = (load latency + hold) / stages.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_stream_bench tma_stream_bench.cu
//   ./tma_stream_bench
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../pathmatfac.jl_b200/csrc/tc_common.cuh"

using namespace pmf::tcx;

__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// mode 0: tensor tiles (2 boxes), mode 1: contiguous 32 KB bulk copies; hold = cycles the consumer keeps a stage
// xload: also load 32 KB of (L2-resident) operand tiles per tile into a second ring, as the data pass does
template <int S>
__global__ void __launch_bounds__(128, 1)
stream_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmX, const float* blocked,
              int n_jt, int n_it, int mode, int hold, int xload, unsigned long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t AG = base, XB = base + S * 32768, BARS = XB + 2 * 32768;
    auto full = [&](int s) { return BARS + 8u * s; };
    auto empty = [&](int s) { return BARS + 8u * (S + s); };
    auto xfull = [&](int s) { return BARS + 8u * (2 * S + s); };
    auto xempty = [&](int s) { return BARS + 8u * (2 * S + 2 + s); };
    if (threadIdx.x == 0) {
        for (int s = 0; s < 2 * S + 4; ++s) mbar_init(BARS + 8u * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long n_tiles = (long long)n_jt * n_it;
    const long long t0 = n_tiles * blockIdx.x / gridDim.x, t1 = n_tiles * (blockIdx.x + 1) / gridDim.x;
    const long long c0 = clock64();
    if (threadIdx.x == 0) {                 // producer of A
        uint32_t s = 0, ph = 0;
        for (long long t = t0; t < t1; ++t) {
            const int jt = (int)(t / n_it), it = (int)(t - (long long)jt * n_it);
            mbar_wait(empty(s), ph ^ 1);
            mbar_expect_tx(full(s), 32768);
            if (mode == 0 || mode == 2) {
                tma_load_2d(AG + s * 32768, &tmA, full(s), it * 64, jt * 128);
                tma_load_2d(AG + s * 32768 + 16384, &tmA, full(s), it * 64 + 32, jt * 128);
                if (mode == 2 && it + 4 < n_it) {     // L2 prefetch four tiles ahead
                    tma_prefetch_2d(&tmA, it * 64 + 256, jt * 128);
                    tma_prefetch_2d(&tmA, it * 64 + 288, jt * 128);
                }
            } else {
                bulk_load_1d(AG + s * 32768, blocked + (size_t)t * 8192, 32768, full(s));
            }
            if (++s == S) { s = 0; ph ^= 1; }
        }
    } else if (threadIdx.x >= 32 && threadIdx.x < 32 + S) {         // one consumer per stage: the holds of the stages overlap
        const uint32_t s = threadIdx.x - 32;
        uint32_t ph = 0;
        for (long long t = t0 + s; t < t1; t += S) {
            mbar_wait(full(s), ph);
            if (hold) { const long long c = clock64(); while (clock64() - c < hold) {} }
            mbar_arrive(empty(s));
            ph ^= 1;
        }
    } else if (threadIdx.x == 64 && xload) {   // producer of the operand tiles (L2 hits: 10 000 x 64 floats = 2.5 MB)
        uint32_t s = 0, ph = 0;
        for (long long t = t0; t < t1; ++t) {
            const int it = (int)(t % n_it);
            mbar_wait(xempty(s), ph ^ 1);
            mbar_expect_tx(xfull(s), 32768);
            for (int k = 0; k < 2; ++k) {
                tma_load_2d(XB + s * 32768 + k * 8192, &tmX, xfull(s), 32 * k, it * 64);
                tma_load_2d(XB + s * 32768 + 16384 + k * 8192, &tmX, xfull(s), 32 * k, it * 64);
            }
            if (++s == 2) { s = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 96 && xload) {
        uint32_t s = 0, ph = 0;
        for (long long t = t0; t < t1; ++t) {
            mbar_wait(xfull(s), ph);
            mbar_arrive(xempty(s));
            if (++s == 2) { s = 0; ph ^= 1; }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(clock64() - c0);
}

template <int S>
float run(const CUtensorMap& tmA, const CUtensorMap& tmX, const float* blocked, int n_jt, int n_it, int mode, int hold,
          int xload, unsigned long long* cyc) {
    const int smem = S * 32768 + 2 * 32768 + 1024 + 256;
    cudaFuncSetAttribute(stream_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        stream_kernel<S><<<148, 128, smem>>>(tmA, tmX, blocked, n_jt, n_it, mode, hold, xload, cyc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
    return best;
}

int main() {
    const int M = 10000, N = 30000, lda = 10016;
    const int n_jt = (N + 127) / 128, n_it = (M + 63) / 64;
    float *A, *blocked, *X;
    cudaMalloc(&A, (size_t)(n_jt * 128) * lda * 4);
    cudaMalloc(&blocked, (size_t)n_jt * n_it * 32768);
    cudaMalloc(&X, (size_t)10048 * 64 * 4);
    cudaMemset(A, 0, (size_t)(n_jt * 128) * lda * 4);
    cudaMemset(blocked, 0, (size_t)n_jt * n_it * 32768);
    cudaMemset(X, 0, (size_t)10048 * 64 * 4);
    unsigned long long* cyc;
    cudaMalloc(&cyc, 148 * 8);
    CUtensorMap tmA, tmX;
    if (!encode_map_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, A, lda, (uint64_t)n_jt * 128, (uint64_t)lda * 4, 32, 128,
                       CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, true) ||
        !encode_map_2d(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, X, 64, 10048, 64 * 4, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B)) {
        printf("tensor map encoding failed\n");
        return 1;
    }
    const double bytes = (double)n_jt * n_it * 32768;
    printf("A stream of %d x %d tiles (%.3f GB); times are the best of 4 launches\n", n_jt, n_it, bytes / 1e9);
    for (int mode = 0; mode < 3; ++mode)
        for (int xload = 1; xload < 2; ++xload)
            for (int hold : {0, 2000, 3000, 4000, 5000}) {
                float t2 = run<2>(tmA, tmX, blocked, n_jt, n_it, mode, hold, xload, cyc);
                float t3 = run<3>(tmA, tmX, blocked, n_jt, n_it, mode, hold, xload, cyc);
                float t4 = run<4>(tmA, tmX, blocked, n_jt, n_it, mode, hold, xload, cyc);
                float t5 = run<5>(tmA, tmX, blocked, n_jt, n_it, mode, hold, xload, cyc);
                printf("%s xload=%d hold=%4d cycles: S=2 %.3f ms (%.0f GB/s) | S=3 %.3f ms (%.0f GB/s) | S=4 %.3f ms (%.0f GB/s) | S=5 %.3f ms (%.0f GB/s)\n",
                       mode == 0 ? "tensor tiles [N][lda]  " : mode == 1 ? "bulk 32 KB tile-blocked" : "tensor tiles + L2 pf 4 ", xload, hold, t2, bytes / t2 / 1e6, t3,
                       bytes / t3 / 1e6, t4, bytes / t4 / 1e6, t5, bytes / t5 / 1e6);
            }
    return 0;
}
