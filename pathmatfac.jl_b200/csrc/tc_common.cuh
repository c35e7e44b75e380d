// PTX wrappers, descriptors and TMEM access macros shared by the tcgen05 kernels (fused_tc.cu: K <= 64 fused data pass;
// wide_tc.cu: K > 64 contraction kernels).  Everything here is inline device code.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace pmf {
namespace tcx {

// ---- PTX wrappers --------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Arrive that publishes nothing: the arriving thread only reports that it has finished READING (TMEM
// accumulators, after tcgen05.wait::ld + fence).  No release fence, so it does not wait for the thread's
// outstanding global loads / stores the way the default (release) arrive does.
__device__ __forceinline__ void mbar_arrive_relaxed(uint32_t bar) {
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
// Four barrier probes issued back to back (their shared-memory round trips overlap); bit k of the result
// is set when barrier k has completed the phase with parity par_k.
__device__ __forceinline__ uint32_t mbar_probe4(uint32_t b0, uint32_t p0, uint32_t b1, uint32_t p1, uint32_t b2,
                                                uint32_t p2, uint32_t b3, uint32_t p3) {
    uint32_t m;
    asm volatile(
        "{\n\t.reg .pred q0, q1, q2, q3;\n\t.reg .u32 t0, t1, t2, t3;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 q0, [%1], %2;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 q1, [%3], %4;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 q2, [%5], %6;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 q3, [%7], %8;\n\t"
        "selp.u32 t0, 1, 0, q0;\n\tselp.u32 t1, 2, 0, q1;\n\tselp.u32 t2, 4, 0, q2;\n\tselp.u32 t3, 8, 0, q3;\n\t"
        "or.b32 t0, t0, t1;\n\tor.b32 t2, t2, t3;\n\tor.b32 %0, t0, t2;\n\t}"
        : "=r"(m) : "r"(b0), "r"(p0), "r"(b1), "r"(p1), "r"(b2), "r"(p2), "r"(b3), "r"(p3) : "memory");
    return m;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// L2 prefetch of a tile (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
// shared -> global tile, element-wise f32 add performed by the memory system (bulk async-group completion)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// tcgen05.mma / commit are issued by ONE thread of the MMA warp (the whole role runs under a single
// elect), so no per-instruction election or vote is needed around them.
// D[tmem] (+)= A[tmem] * B[smem]   (kind::tf32, cta_group::1)
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// same with 16-bit operands (kind::f16, here BF16 x BF16 -> F32, K = 16 per instruction)
__device__ __forceinline__ void mma_ts_f16(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint32_t bar) { tc_commit(bar); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\tselp.u32 %0, 1, 0, e;\n\t}" : "=r"(pred));
    return pred != 0;
}

// Shared-memory operand descriptors, 128-byte swizzle (rows of 32 floats, 8-row groups 1024 B apart).
// K-major: the contraction runs along the 128-byte rows; SBO = 1024 separates 8-row groups of M / N.
__device__ __forceinline__ uint64_t umma_desc_k(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// MN-major FP32/TF32 operand: M / N runs along the 128-byte rows, the contraction across rows.  The only
// layout the tensor core accepts here is "128B swizzle with 32-byte atoms" (layout type 1): 32-byte
// chunk index ^= row & 3, i.e. a 4-row x 128 B atom (TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).
// LBO = distance between consecutive 32-element chunks of M / N (one TMA box), SBO = 512 between the
// two 4-row groups of one k-step (8 rows).
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)(lbo_bytes >> 4) << 16) | (32ull << 32) | (1ull << 46) |
           (1ull << 61);
}
// instruction descriptor: TF32 x TF32 -> F32; bit 15 / 16 = A / B operand is MN-major
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

#define TMEM_LD32(taddr, r)                                                                                       \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                        \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"  \
                 "%26,%27,%28,%29,%30,%31}, [%32];"                                                               \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),  \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),        \
                   "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),      \
                   "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),      \
                   "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                          \
                 : "r"(taddr))
// 16 lanes x 256 bits, 8 times along the columns: a 16 x 64 block of 32-bit words spread over the 32 threads of the warp
// (thread t: rows t/4 and t/4 + 8; per 8-column atom j the registers r[4j], r[4j+1] = row t/4, columns 8j + 2(t%4), +1
// and r[4j+2], r[4j+3] = row t/4 + 8, same columns)
#define TMEM_LD_16x256b_x8(taddr, r)                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 "                                                        \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"  \
                 "%26,%27,%28,%29,%30,%31}, [%32];"                                                               \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),  \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),        \
                   "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),      \
                   "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),      \
                   "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                          \
                 : "r"(taddr))
#define TMEM_ST32(taddr, r)                                                                                       \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                  \
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26," \
                 "%27,%28,%29,%30,%31,%32};"                                                                      \
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),        \
                 "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),      \
                 "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),   \
                 "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),   \
                 "r"(r[31]) : "memory")

#define TMEM_ST8(taddr, r)                                                                                        \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"                         \
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),        \
                 "r"(r[7]) : "memory")

#define TMEM_LD8(taddr, r)                                                                                        \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                          \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])   \
                 : "r"(taddr))
#define TMEM_LD16(taddr, r)                                                                                       \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                        \
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                                  \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),  \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),        \
                   "=r"(r[15])                                                                                    \
                 : "r"(taddr))
#define TMEM_ST16(taddr, r)                                                                                       \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "                                                  \
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"                                        \
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]),        \
                 "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),      \
                 "r"(r[15]) : "memory")

// two floats -> packed BF16 pair, `lo` in bits 0..15 (the even k of a 16-bit tensor-core operand)
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ uint32_t rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

// TF32 operand rounding for values the tensor core will truncate anyway: adding half a TF32 ulp to the
// bit pattern makes that truncation a round-to-nearest (ties away); one integer add per entry.
__device__ __forceinline__ uint32_t rn_bits(float x) { return __float_as_uint(x) + 0x1000u; }

// single-instruction SFU forms (flush-to-zero: no denormal range fix-up code around the MUFU)
__device__ __forceinline__ float ex2_fast(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_fast(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_fast(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// all-ones when the datum is observed (finite), zero when it is missing: ANDed into results that are
// NaN for a missing datum, so the mask costs no branch
__device__ __forceinline__ uint32_t obs_mask(float a) { return fabsf(a) < INFINITY ? 0xffffffffu : 0u; }
__device__ __forceinline__ float and_mask(float x, uint32_t m) { return __uint_as_float(__float_as_uint(x) & m); }


// ---- host side: tensor-map encoding (driver entry point fetched at run time: libpmf does not link libcuda) ----------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
    return fn;
}
// 2-D tensor [rows][cols] (cols contiguous, row pitch in bytes), box = box_rows x box_cols elements
static inline bool encode_map_2d(CUtensorMap* m, CUtensorMapDataType dt, const void* base, uint64_t cols, uint64_t rows,
                                 uint64_t pitch_bytes, uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle sw,
                                 bool nan_fill = false) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {pitch_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
               CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               nan_fill ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace tcx
}  // namespace pmf
