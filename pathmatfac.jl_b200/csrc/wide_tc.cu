// Data pass for 64 < K <= 256 on the 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// K is the number of pathway factors (src/model.jl:122, K = length(feature_graphs)); at K = 128 / 256 the three
// contractions of the pass (Z = X'Y, dX = Y G', dY = X G) are tensor-bound and the operand
// tiles of a single fused tile pipeline (fused_tc.cu) no longer fit TMEM / shared memory.  The pass is therefore three
// GEMM-shaped kernels around ONE M x N scratch matrix G' (4 HBM bytes per entry written once, read twice):
//
//   zlink_kernel      Z[j,i] = sum_k Y[j,k] X[i,k] from a two-term BF16 split of both operands, h = bf16(v),
//                     l = bf16(v - h): Yh Xh + Yh Xl + Yl Xh, three BF16 contractions with FP32 accumulation in TMEM
//                     (operand error 2^-17; measured |dZ| = 4e-6 rms(Z), against 7e-7 for the TF32 + BF16-correction
//                     form of fused_tc.cu -- at half the operand bytes per factor and 3/4 of its tensor time, which
//                     matters here because the K-loop re-streams both operands from L2 for every tile and the L2 -> SM
//                     fabric, not the tensor pipe, bounds it).  K-loop over 64-wide slabs, operands K-major through a
//                     TMA-fed ring; accumulator tile 128 features x 128 samples, four of them in TMEM.  Epilogue (16
//                     warps, lane = feature): ColScale / ColShift, noise-model loss and dloss/dz with the NaN mask
//                     (src/layers.jl:9-90), column sums for dmu / dlogsigma; the data tile comes in and
//                     G' = w_j sigma_j dloss/dz goes out through shared memory by TMA.
//   grad_gemm_kernel  <A_MN = false>  dY[j,:] += sum_i G'[j,i] Xh[i,:]   A = G' tile K-major, B = Xh rows MN-major
//                     <A_MN = true >  dX[i,:] += sum_j G'[j,i] Yh[j,:]   A = G' tile MN-major (no transposition), B = Yh
//                     two 128-row output tiles x K columns accumulate in TMEM over a contiguous range of 32-deep
//                     contraction steps; the accumulators leave once per item as 128-bit REDs.
//
// Operand scratch (prep_wide_kernel, every epoch): Ph = rna_tf32(P) [rows][Kq] FP32 (gradient contractions) and two BF16
// planes Pb = [bf16(P) | bf16(P - bf16(P))], each [rows][Kq] (Z contraction); Kq = roundup(K, 64), zero padded.
// Single-pass TF32 with round-to-nearest operands for the two gradient contractions (as in fused_tc.cu).
// Batch shift / scale layers are not served here (the FP32 kernel runs those models when K > 64).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "pmf_epilogue.cuh"
#include "pmf_internal.h"
#include "tc_common.cuh"

namespace pmf {

namespace {

using namespace tcx;

// ---- work distribution --------------------------------------------------------------------------------------
// Units (outer, inner), flattened outer-major, are cut into gridDim.x contiguous ranges of equal cost; a range is
// walked as ITEMS = maximal runs inside one outer index (per-item state: column constants, accumulators).
// `cum`: cumulative cost per outer index ([n_outer + 1]) or null for uniform cost.
struct RangeIter {
    long long t, t_end;
    int n_inner, outer, in0, in1;
    static __device__ __forceinline__ long long cut(const int32_t* cum, int n_outer, int n_inner, unsigned num, unsigned den) {
        const long long total = (long long)n_outer * n_inner;
        if (num >= den) return total;
        if (cum == nullptr) return total * num / den;
        const long long W = (long long)cum[n_outer] * n_inner;
        const long long x = W * num / den;
        int lo = 0, hi = n_outer;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if ((long long)cum[mid] * n_inner <= x) lo = mid; else hi = mid;
        }
        const long long w = cum[lo + 1] - cum[lo];
        long long in = (x - (long long)cum[lo] * n_inner) / (w > 0 ? w : 1);
        if (in > n_inner) in = n_inner;
        return (long long)lo * n_inner + in;
    }
    __device__ __forceinline__ RangeIter(const int32_t* cum, int n_outer, int n_inner_) {
        t = cut(cum, n_outer, n_inner_, blockIdx.x, gridDim.x);
        t_end = cut(cum, n_outer, n_inner_, blockIdx.x + 1, gridDim.x);
        n_inner = n_inner_;
        outer = in0 = in1 = 0;
    }
    __device__ __forceinline__ bool next() {
        if (t >= t_end) return false;
        outer = (int)(t / n_inner);
        in0 = (int)(t - (long long)outer * n_inner);
        const long long len = min((long long)(n_inner - in0), t_end - t);
        in1 = in0 + (int)len;
        t += len;
        return true;
    }
};

struct Ring {
    uint32_t s = 0, ph = 0;
    __device__ __forceinline__ void next(uint32_t n) { if (++s == n) { s = 0; ph ^= 1u; } }
};

// D[tmem] (+)= A[smem] * B[smem], 16-bit operands (kind::f16; here BF16 x BF16 -> F32, K = 16 per instruction)
__device__ __forceinline__ void mma_ss_f16(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// shared -> global tile store (bulk async-group completion); out-of-bounds parts of the box are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}

__device__ __noinline__ float2 noise_eval_slow_w(int dist, float z, float a, const float* __restrict__ th_range,
                                                  float ord_eps, float margin) {
    if (!is_observed(a)) return make_float2(0.f, 0.f);
    const float4 th4 = __ldg(reinterpret_cast<const float4*>(th_range));
    const float th[4] = {th4.x, th4.y, th4.z, th4.w};
    float l, g;
    noise_eval(dist, z, a, th, ord_eps, margin, l, g);
    return make_float2(l, g);
}

__device__ __noinline__ float2 threshold_grads_slow_w(int dist, float z, float a, const float* __restrict__ th_range,
                                                      float ord_eps, float margin) {
    if (!is_observed(a)) return make_float2(0.f, 0.f);
    const float4 th4 = __ldg(reinterpret_cast<const float4*>(th_range));
    const float th[4] = {th4.x, th4.y, th4.z, th4.w};
    float g1, g2;
    noise_threshold_grads(dist, z, a, th, ord_eps, margin, g1, g2);
    return make_float2(g1, g2);
}

// ================================================================================================================
// zlink: Z contraction + link epilogue
// ================================================================================================================
// Tile = 128 features x 128 samples, accumulators in TMEM (4 x 128 columns: the K-loop runs up to three tiles ahead
// of the epilogue).  The data take the TMA path in and out: a tile's 128 x 128 block of A arrives as two 64-sample
// sub-tiles (two 32-sample boxes each, NaN fill beyond the matrix) in a 3-deep shared-memory ring, the 16 epilogue
// warps turn a sub-tile into G' IN PLACE (128-bit accesses, conflict free under the 128-byte swizzle) and a TMA store
// returns it.  Measured on B200 at 10 000 x 50 000, K = 128: per-lane 256-bit global loads / stores of the same data
// (a lane per feature row, 32-byte sectors 40 KB apart) cost 0.77 ms on top of the 0.55 ms K-loop; the bulk path
// moves the same bytes at HBM speed behind the tensor work.
constexpr int ZBJ = 128, ZBI = 128, ZS = 2, ZSZ = 4, ZSA = 3;
constexpr uint32_t Z_YH = 16384, Z_YB = 16384, Z_XH = 16384, Z_XB = 16384;
constexpr uint32_t Z_OFF_YB = Z_YH, Z_OFF_XH = Z_YH + Z_YB, Z_OFF_XB = Z_YH + Z_YB + Z_XH;
constexpr uint32_t Z_STAGE = Z_YH + Z_YB + Z_XH + Z_XB;                      // 64 KB per 64-factor slab: Yh | Yl | Xh | Xl (BF16)
constexpr uint32_t Z_AG = 32768;                                            // 128 features x 64 samples of A / G'
constexpr int Z_NEPI = 16, Z_W_TMA = Z_NEPI, Z_W_MMA = Z_NEPI + 1, Z_W_TMA_A = Z_NEPI + 2, Z_W_GST = Z_NEPI + 3,
              Z_NTHREADS = 32 * (Z_NEPI + 4);
constexpr uint32_t Z_SMEM_DATA = ZS * Z_STAGE + ZSA * Z_AG;
constexpr uint32_t Z_SMEM = Z_SMEM_DATA + 1024 + 256;
enum ZBar { ZB_FULL = 0, ZB_EMPTY = ZB_FULL + ZS, ZB_ZFULL = ZB_EMPTY + ZS, ZB_ZEMPTY = ZB_ZFULL + ZSZ,
            ZB_AG_FULL = ZB_ZEMPTY + ZSZ, ZB_AG_EMPTY = ZB_AG_FULL + ZSA, ZB_G_READY = ZB_AG_EMPTY + ZSA,
            ZB_COUNT = ZB_G_READY + ZSA };

struct WideParams {
    DataPassParams dp;
    int n_jt, n_it;      // 128-feature tiles, 128-sample tiles
    int nks;             // 64-factor slabs = Kq / 64
    int flags;           // PMF_WIDE_FLAGS experiments (results wrong): 1 no data loads, 2 no G' stores
};

__global__ void __launch_bounds__(Z_NTHREADS, 1)
zlink_kernel(const __grid_constant__ CUtensorMap tmYh, const __grid_constant__ CUtensorMap tmYb,
             const __grid_constant__ CUtensorMap tmXh, const __grid_constant__ CUtensorMap tmXb,
             const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmG, const WideParams p) {
    const DataPassParams& dp = p.dp;
    if (dp.stop_flag != nullptr && *dp.stop_flag != 0) return;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t AG = base + ZS * Z_STAGE;
    uint8_t* ag_ptr0 = gbase + ZS * Z_STAGE;
    const uint32_t BARS = base + Z_SMEM_DATA;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + Z_SMEM_DATA + 8 * ZB_COUNT);
    __shared__ double red_smem[Z_NEPI];
    auto bar = [&](int b) { return BARS + 8u * (uint32_t)b; };
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int b = 0; b < ZB_COUNT; ++b) {
            const bool per_warp = (b >= ZB_ZEMPTY && b < ZB_ZEMPTY + ZSZ) || (b >= ZB_G_READY && b < ZB_G_READY + ZSA);
            mbar_init(bar(b), per_warp ? (uint32_t)Z_NEPI : 1u);     // one arrive per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == Z_W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tmem_slot;

    if (warp == Z_W_TMA) {
        // ================================ operand producer (L2) ====================================
        if (lane == 0) {
            Ring r;
            for (RangeIter itx(dp.tc_cost_cum, p.n_jt, p.n_it); itx.next();) {
                const int j0 = itx.outer * ZBJ;
                for (int it = itx.in0; it < itx.in1; ++it) {
                    const int i0 = it * ZBI;
                    for (int ks = 0; ks < p.nks; ++ks, r.next(ZS)) {
                        mbar_wait(bar(ZB_EMPTY + r.s), r.ph ^ 1);
                        mbar_expect_tx(bar(ZB_FULL + r.s), Z_STAGE);
                        const uint32_t st = base + r.s * Z_STAGE;
                        tma_load_2d(st, &tmYh, bar(ZB_FULL + r.s), 64 * ks, j0);
                        tma_load_2d(st + Z_OFF_XH, &tmXh, bar(ZB_FULL + r.s), 64 * ks, i0);
                        tma_load_2d(st + Z_OFF_YB, &tmYb, bar(ZB_FULL + r.s), 64 * ks, j0);
                        tma_load_2d(st + Z_OFF_XB, &tmXb, bar(ZB_FULL + r.s), 64 * ks, i0);
                    }
                }
            }
        }
    } else if (warp == Z_W_TMA_A) {
        // ================================ data producer (HBM stream) ===============================
        if (lane == 0) {
            Ring r;
            for (RangeIter itx(dp.tc_cost_cum, p.n_jt, p.n_it); itx.next();) {
                const int j0 = itx.outer * ZBJ;
                for (int it = itx.in0; it < itx.in1; ++it) {
                    for (int h = 0; h < 2; ++h, r.next(ZSA)) {
                        const int i0 = it * ZBI + 64 * h;
                        mbar_wait(bar(ZB_AG_EMPTY + r.s), r.ph ^ 1);
                        if (p.flags & 1) { mbar_arrive(bar(ZB_AG_FULL + r.s)); continue; }
                        mbar_expect_tx(bar(ZB_AG_FULL + r.s), Z_AG);
                        for (int q = 0; q < 2; ++q)
                            tma_load_2d(AG + r.s * Z_AG + q * 16384, &tmA, bar(ZB_AG_FULL + r.s), i0 + 32 * q, j0);
                    }
                }
            }
        }
    } else if (warp == Z_W_GST) {
        // ================================ G' store issuer ==========================================
        // One TMA store per 32-sample box; a buffer returns to the data producer once the store engine has READ it.
        // One store group stays in flight: buffer u - 1 is released after the stores of u have been issued.
        if (lane == 0) {
            Ring r;
            int prev = -1;
            for (RangeIter itx(dp.tc_cost_cum, p.n_jt, p.n_it); itx.next();) {
                const int j0 = itx.outer * ZBJ;
                for (int it = itx.in0; it < itx.in1; ++it) {
                    for (int h = 0; h < 2; ++h, r.next(ZSA)) {
                        const int i0 = it * ZBI + 64 * h;
                        mbar_wait(bar(ZB_G_READY + r.s), r.ph);
                        if (!(p.flags & 2)) {
                            for (int q = 0; q < 2; ++q)
                                if (i0 + 32 * q < dp.lda) tma_store_2d(&tmG, AG + r.s * Z_AG + q * 16384, i0 + 32 * q, j0);
                        }
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        if (prev >= 0) mbar_arrive(bar(ZB_AG_EMPTY + prev));
                        prev = (int)r.s;
                    }
                }
            }
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (prev >= 0) mbar_arrive(bar(ZB_AG_EMPTY + prev));
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");       // every store has been performed
        }
    } else if (warp == Z_W_MMA) {
        if (elect_one()) {
            // BF16 x BF16 -> F32 (kind::f16), both operands K-major, K = 16 per instruction
            const uint32_t id_zb = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(ZBI >> 3) << 17) | ((uint32_t)(ZBJ >> 4) << 24);
            Ring r, rz;
            for (RangeIter itx(dp.tc_cost_cum, p.n_jt, p.n_it); itx.next();) {
                for (int it = itx.in0; it < itx.in1; ++it, rz.next(ZSZ)) {
                    mbar_wait(bar(ZB_ZEMPTY + rz.s), rz.ph ^ 1u);
                    tc_fence_after();
                    const uint32_t zt = tm + ZBI * rz.s;
                    for (int ks = 0; ks < p.nks; ++ks, r.next(ZS)) {
                        mbar_wait(bar(ZB_FULL + r.s), r.ph);
                        tc_fence_after();
                        const uint32_t st = base + r.s * Z_STAGE;
                        const uint64_t yh = umma_desc_k(st), yl = umma_desc_k(st + Z_OFF_YB);
                        const uint64_t xh = umma_desc_k(st + Z_OFF_XH), xl = umma_desc_k(st + Z_OFF_XB);
#pragma unroll
                        for (int s = 0; s < 4; ++s)       // Yh Xh
                            mma_ss_f16(zt, yh + (uint64_t)(2 * s), xh + (uint64_t)(2 * s), id_zb, (ks > 0 || s > 0) ? 1u : 0u);
#pragma unroll
                        for (int s = 0; s < 4; ++s)       // Yh Xl
                            mma_ss_f16(zt, yh + (uint64_t)(2 * s), xl + (uint64_t)(2 * s), id_zb, 1u);
#pragma unroll
                        for (int s = 0; s < 4; ++s)       // Yl Xh
                            mma_ss_f16(zt, yl + (uint64_t)(2 * s), xh + (uint64_t)(2 * s), id_zb, 1u);
                        tc_commit(bar(ZB_EMPTY + r.s));
                    }
                    tc_commit(bar(ZB_ZFULL + rz.s));
                }
            }
        }
        __syncwarp();
    } else {
        // ================================ epilogue warps ===========================================
        // warp = (TMEM lane quarter, 16-sample chunk of the 64-sample sub-tile); lane = feature row
        const int quarter = warp & 3, c16 = warp >> 2;
        const int lrow = 32 * quarter + lane;
        const uint32_t lane_addr = ((uint32_t)(32 * quarter)) << 16;
        // this thread's 64 bytes of its row in box (c16 >> 1): logical 16-byte chunks 4 (c16 & 1) .. + 3 of the 128-byte
        // row, physical chunk = logical ^ (row & 7)  (128-byte swizzle)
        const uint32_t row_off = (uint32_t)(c16 >> 1) * 16384u + (uint32_t)lrow * 128u;
        auto chunk_off = [&](int v) { return row_off + (uint32_t)(((4 * (c16 & 1) + v) ^ (lrow & 7)) << 4); };
        double loss_d = 0.0;
        Ring ra, rz;
        for (RangeIter itx(dp.tc_cost_cum, p.n_jt, p.n_it); itx.next();) {
            const int j = itx.outer * ZBJ + lrow;
            const bool jok = j < dp.N;
            const int jj = jok ? j : dp.N - 1;           // padding rows follow the last column (same branch; their data are NaN)
            const float sigma = __expf(__ldg(dp.logsigma + jj));
            const float muj = __ldg(dp.mu + jj);
            const float wj = jok ? __ldg(dp.weight + jj) : 0.f;
            const int ci = __ldg(dp.colinfo + jj);
            const int dist = ci & 0xff;
            const float* th_range = dp.thresholds + 4 * (ci >> 8);
            const float gscale = sigma * wj;
            float dmu_acc = 0.f, loss_acc = 0.f;
            for (int it = itx.in0; it < itx.in1; ++it, rz.next(ZSZ)) {
                mbar_wait(bar(ZB_ZFULL + rz.s), rz.ph);
                tc_fence_after();
#pragma unroll 1
                for (int h = 0; h < 2; ++h, ra.next(ZSA)) {
                    uint32_t z[16];
                    float a[16];
                    TMEM_LD16(tm + lane_addr + ZBI * rz.s + 64 * h + 16 * c16, z);
                    mbar_wait(bar(ZB_AG_FULL + ra.s), ra.ph);
                    uint8_t* abox = ag_ptr0 + ra.s * Z_AG;
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const float4 a4 = *reinterpret_cast<const float4*>(abox + chunk_off(v));
                        a[4 * v] = a4.x; a[4 * v + 1] = a4.y; a[4 * v + 2] = a4.z; a[4 * v + 3] = a4.w;
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (h == 1) {
                        // both halves of the accumulator have been read by this warp: the K-loop may reuse it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_relaxed(bar(ZB_ZEMPTY + rz.s));
                    }
                    if (dist == DIST_NORMAL) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            float d = fmaf(__uint_as_float(z[e]), sigma, muj) - a[e];
                            d = fabsf(a[e]) < INFINITY ? d : 0.f;
                            loss_acc = fmaf(d, d, loss_acc);
                            dmu_acc += d;
                            z[e] = rn_bits(d * gscale);
                        }
                    } else if (dist == DIST_BERNOULLI) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const uint32_t m = obs_mask(a[e]);
                            float z4 = fmaf(__uint_as_float(z[e]), sigma, muj);
                            float ex = ex2_fast(fabsf(z4) * -1.4426950408889634f);
                            float w = 1.0f + ex;
                            float r = rcp_fast(w);
                            float sg = z4 >= 0.f ? r : ex * r;
                            float l = fmaf(-a[e], z4, fmaf(lg2_fast(w), 0.6931471805599453f, fmaxf(z4, 0.f)));
                            float gv = and_mask(sg - a[e], m);
                            loss_acc = fmaf(2.f, and_mask(l, m), loss_acc);
                            dmu_acc += gv;
                            z[e] = rn_bits(gv * gscale);
                        }
                    } else if (dist == DIST_POISSON) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const uint32_t m = obs_mask(a[e]);
                            float z4 = fmaf(__uint_as_float(z[e]), sigma, muj);
                            float ez = ex2_fast(z4 * 1.4426950408889634f);
                            float gv = and_mask(ez - a[e], m);
                            float l = and_mask(fmaf(-a[e], z4, ez), m);
                            loss_acc = fmaf(2.f, l, loss_acc);
                            dmu_acc += gv;
                            z[e] = rn_bits(gv * gscale);
                        }
                    } else {
                        float t1s = 0.f, t2s = 0.f;      // interior-threshold gradients (update_noise_models)
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            float z4 = fmaf(__uint_as_float(z[e]), sigma, muj);
                            float2 lg = noise_eval_slow_w(dist, z4, a[e], th_range, dp.ordinal_eps, dp.hinge_margin);
                            if (dp.dthr != nullptr && is_ordinal(dist)) {
                                float2 tg = threshold_grads_slow_w(dist, z4, a[e], th_range, dp.ordinal_eps, dp.hinge_margin);
                                t1s += tg.x; t2s += tg.y;
                            }
                            loss_acc = fmaf(2.f, lg.x, loss_acc);
                            dmu_acc += lg.y;
                            z[e] = rn_bits(lg.y * gscale);
                        }
                        if (dp.dthr != nullptr && is_ordinal(dist))
                            add_threshold_grads(dp.dthr, ci >> 8, t1s * wj, t2s * wj);
                    }
                    // G' over the data values this thread just read (same addresses), then hand the buffer to the store
#pragma unroll
                    for (int v = 0; v < 4; ++v)
                        *reinterpret_cast<uint4*>(abox + chunk_off(v)) = make_uint4(z[4 * v], z[4 * v + 1], z[4 * v + 2], z[4 * v + 3]);
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar(ZB_G_READY + ra.s));
                }
            }
            loss_d += (double)(loss_acc * wj);
            // dmu_j = w_j sum_i g ; dlogsigma_j = sigma_j w_j sum_i g (the reference's ColScale quirk, src/layers.jl:39-44)
            if (jok) {
                atomicAdd(dp.dmu + j, dmu_acc * wj);
                atomicAdd(dp.dlogsigma + j, dmu_acc * wj * sigma);
            }
        }
        loss_d *= 0.5;
        for (int o = 16; o > 0; o >>= 1) loss_d += __shfl_xor_sync(0xffffffffu, loss_d, o);
        if (lane == 0) red_smem[warp] = loss_d;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < Z_NEPI; ++w) t += red_smem[w];
        atomicAdd(dp.scalars + SC_DATA, t);
    }
    if (warp == Z_W_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
    }
}

// ================================================================================================================
// gradient contractions
// ================================================================================================================
constexpr int GS = 3;                                  // ring stages
constexpr uint32_t G_A = 32768, G_B = 32768, G_STAGE = G_A + G_B;
constexpr int G_NEPI = 4, G_W_TMA = G_NEPI, G_W_MMA = G_NEPI + 1, G_NTHREADS = 32 * (G_NEPI + 2);
constexpr uint32_t G_SMEM = GS * G_STAGE + 1024 + 256;
enum GBar { GB_FULL = 0, GB_EMPTY = GB_FULL + GS, GB_ACC_FULL = GB_EMPTY + GS, GB_ACC_EMPTY, GB_COUNT };

struct GradParams {
    float* out;          // [rows_pad][Kp] accumulated with REDs
    int rows_pad;        // rows of `out` (multiple of 128)
    int Kp, Kc, nkb;     // row pitch, accumulator width (multiple of 16), 32-column boxes of the B operand
    int n_groups;        // 256-row output groups
    int n_steps;         // 32-deep contraction steps
    const int* stop_flag;
};

template <bool A_MN>
__global__ void __launch_bounds__(G_NTHREADS, 1)
grad_gemm_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmB, const GradParams p) {
    if (p.stop_flag != nullptr && *p.stop_flag != 0) return;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t BARS = base + GS * G_STAGE;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + GS * G_STAGE + 8 * GB_COUNT);
    auto bar = [&](int b) { return BARS + 8u * (uint32_t)b; };
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int b = 0; b < GB_COUNT; ++b) mbar_init(bar(b), b == GB_ACC_EMPTY ? (uint32_t)G_NEPI : 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == G_W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tmem_slot;

    if (warp == G_W_TMA) {
        if (lane == 0) {
            Ring r;
            const uint32_t bytes = G_A + (uint32_t)p.nkb * 4096u;
            for (RangeIter itx(nullptr, p.n_groups, p.n_steps); itx.next();) {
                const int r0 = itx.outer * 256;
                for (int c = itx.in0; c < itx.in1; ++c, r.next(GS)) {
                    const int c0 = 32 * c;
                    mbar_wait(bar(GB_EMPTY + r.s), r.ph ^ 1);
                    mbar_expect_tx(bar(GB_FULL + r.s), bytes);
                    const uint32_t st = base + r.s * G_STAGE;
                    if (!A_MN) {
                        // G' rows r0 .. (features), 32 samples: K-major A tiles (contraction = samples, contiguous)
                        for (int t = 0; t < 2; ++t) tma_load_2d(st + t * 16384, &tmG, bar(GB_FULL + r.s), c0, r0 + 128 * t);
                    } else {
                        // 32 feature rows x 128 samples per tile as four 32-sample boxes: MN-major A tiles
                        for (int t = 0; t < 2; ++t)
                            for (int q = 0; q < 4; ++q)
                                tma_load_2d(st + t * 16384 + q * 4096, &tmG, bar(GB_FULL + r.s), r0 + 128 * t + 32 * q, c0);
                    }
                    for (int kb = 0; kb < p.nkb; ++kb) tma_load_2d(st + G_A + kb * 4096, &tmB, bar(GB_FULL + r.s), 32 * kb, c0);
                }
            }
        }
    } else if (warp == G_W_MMA) {
        if (elect_one()) {
            const uint32_t idesc = umma_idesc(128, p.Kc, A_MN, true);
            Ring r;
            uint32_t q = 0;
            for (RangeIter itx(nullptr, p.n_groups, p.n_steps); itx.next(); ++q) {
                mbar_wait(bar(GB_ACC_EMPTY), (q & 1u) ^ 1u);
                tc_fence_after();
                for (int c = itx.in0; c < itx.in1; ++c, r.next(GS)) {
                    mbar_wait(bar(GB_FULL + r.s), r.ph);
                    tc_fence_after();
                    const uint32_t st = base + r.s * G_STAGE;
                    const uint64_t bd = umma_desc_mn(st + G_A, 4096u);
                    const uint32_t acc = c > itx.in0 ? 1u : 0u;
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const uint64_t ad = A_MN ? umma_desc_mn(st + t * 16384, 4096u) : umma_desc_k(st + t * 16384);
#pragma unroll
                        for (int s = 0; s < 4; ++s)
                            mma_ss(tm + (uint32_t)(t * p.Kc), ad + (uint64_t)(A_MN ? 64 * s : 2 * s), bd + (uint64_t)(64 * s), idesc,
                                   s > 0 ? 1u : acc);
                    }
                    tc_commit(bar(GB_EMPTY + r.s));
                }
                tc_commit(bar(GB_ACC_FULL));
            }
        }
        __syncwarp();
    } else {
        const int quarter = warp & 3;
        const uint32_t lane_addr = ((uint32_t)(32 * quarter)) << 16;
        uint32_t q = 0;
        for (RangeIter itx(nullptr, p.n_groups, p.n_steps); itx.next(); ++q) {
            mbar_wait(bar(GB_ACC_FULL), q & 1u);
            tc_fence_after();
            for (int t = 0; t < 2; ++t) {
                const int row = itx.outer * 256 + 128 * t + 32 * quarter + lane;
                float* dst = p.out + (size_t)row * p.Kp;
                for (int c16 = 0; 16 * c16 < p.Kc; ++c16) {
                    uint32_t v[16];
                    TMEM_LD16(tm + lane_addr + (uint32_t)(t * p.Kc + 16 * c16), v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (row < p.rows_pad) {
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (16 * c16 + 4 * u < p.Kp)
                                atomicAdd(reinterpret_cast<float4*>(dst + 16 * c16 + 4 * u),
                                          make_float4(__uint_as_float(v[4 * u]), __uint_as_float(v[4 * u + 1]),
                                                      __uint_as_float(v[4 * u + 2]), __uint_as_float(v[4 * u + 3])));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_relaxed(bar(GB_ACC_EMPTY));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == G_W_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
    }
}

// Operand split of a factor matrix P [rows][Kp] for the wide path (see the file header).  One thread per 4 factors.
__global__ void prep_wide_kernel(const float* __restrict__ P, float* __restrict__ Ph, uint2* __restrict__ Pb, int rows, int Kp,
                                 int Kq, const int* stop_flag) {
    if (stop_flag != nullptr && *stop_flag != 0) return;
    const int q4 = Kq >> 2;
    const size_t n4 = (size_t)rows * q4;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n4; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t row = idx / q4;
        const int k = 4 * (int)(idx - row * q4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < Kp) v = *reinterpret_cast<const float4*>(P + row * Kp + k);
        float4 h;
        h.x = __uint_as_float(rna_tf32(v.x)); h.y = __uint_as_float(rna_tf32(v.y));
        h.z = __uint_as_float(rna_tf32(v.z)); h.w = __uint_as_float(rna_tf32(v.w));
        *reinterpret_cast<float4*>(Ph + row * Kq + k) = h;
        // two-term BF16 split: b = bf16(v) (round to nearest), l = bf16(v - b)
        const uint2 hb = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
        const float bx = __uint_as_float(hb.x << 16), by = __uint_as_float(hb.x & 0xffff0000u);
        const float bz = __uint_as_float(hb.y << 16), bw = __uint_as_float(hb.y & 0xffff0000u);
        const uint2 lb = make_uint2(pack_bf16(v.x - bx, v.y - by), pack_bf16(v.z - bz, v.w - bw));
        Pb[idx] = hb;                               // plane 0: [rows][Kq] BF16
        Pb[n4 + idx] = lb;                          // plane 1
    }
}

}  // namespace

bool wide_supported(const DataPassParams& p) {
    return p.Kp > 64 && p.Kp <= 256 && p.n_batch_views == 0 && p.col_ssq == nullptr;
}

size_t wide_scratch_floats(int rows_pad, int Kp) { return (size_t)rows_pad * (size_t)((Kp + 63) / 64 * 64); }

cudaError_t launch_data_pass_wide(const DataPassParams& dp, const WideScratch& ws, int precision, cudaStream_t s, int n_sms,
                                  int* n_launches) {
    if (!wide_supported(dp)) return cudaErrorInvalidValue;
    const int Kq = (dp.Kp + 63) / 64 * 64;
    const int Kc = (dp.Kp + 15) / 16 * 16;
    int launched = 0;
    // 1. operand split of X and Y
    {
        const size_t nx = (size_t)dp.Mp * (Kq >> 2), ny = (size_t)dp.Np * (Kq >> 2);
        auto blocks = [](size_t n) { size_t b = (n + 255) / 256; return (unsigned)(b < 1184 ? (b ? b : 1) : 1184); };
        prep_wide_kernel<<<blocks(nx), 256, 0, s>>>(dp.X, ws.Xh, reinterpret_cast<uint2*>(ws.Xb), dp.Mp, dp.Kp, Kq, dp.stop_flag);
        prep_wide_kernel<<<blocks(ny), 256, 0, s>>>(dp.Y, ws.Yh, reinterpret_cast<uint2*>(ws.Yb), dp.Np, dp.Kp, Kq, dp.stop_flag);
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        launched += 2;
    }
    const CUtensorMapDataType F32 = CU_TENSOR_MAP_DATA_TYPE_FLOAT32, BF16 = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const CUtensorMapSwizzle SW = CU_TENSOR_MAP_SWIZZLE_128B, SWA = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    // 2. Z + link
    {
        CUtensorMap tmYh, tmYb, tmXh, tmXb, tmA, tmG;
        // BF16 planes: hi at the base, lo one plane (rows x Kq elements) further
        const uint16_t* yb = static_cast<const uint16_t*>(ws.Yb);
        const uint16_t* xb = static_cast<const uint16_t*>(ws.Xb);
        bool ok = encode_map_2d(&tmYh, BF16, yb, Kq, dp.Np, (uint64_t)Kq * 2, 64, ZBJ, SW) &&
                  encode_map_2d(&tmYb, BF16, yb + (size_t)dp.Np * Kq, Kq, dp.Np, (uint64_t)Kq * 2, 64, ZBJ, SW) &&
                  encode_map_2d(&tmXh, BF16, xb, Kq, dp.Mp, (uint64_t)Kq * 2, 64, ZBI, SW) &&
                  encode_map_2d(&tmXb, BF16, xb + (size_t)dp.Mp * Kq, Kq, dp.Mp, (uint64_t)Kq * 2, 64, ZBI, SW) &&
                  encode_map_2d(&tmA, F32, dp.A, dp.lda, dp.N, (uint64_t)dp.lda * 4, 32, ZBJ, SW, /*nan_fill=*/true) &&
                  encode_map_2d(&tmG, F32, ws.G, dp.lda, dp.N, (uint64_t)dp.lda * 4, 32, ZBJ, SW);
        if (!ok) return cudaErrorUnknown;
        WideParams p;
        p.dp = dp;
        p.n_jt = (dp.N + ZBJ - 1) / ZBJ;
        p.n_it = (dp.M + ZBI - 1) / ZBI;
        p.nks = Kq / 64;
        { const char* f = getenv("PMF_WIDE_FLAGS"); p.flags = f ? atoi(f) : 0; }
        cudaError_t e = cudaFuncSetAttribute(zlink_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Z_SMEM);
        if (e != cudaSuccess) return e;
        const long long n_tiles = (long long)p.n_jt * p.n_it;
        const int grid = n_tiles < n_sms ? (int)n_tiles : n_sms;
        zlink_kernel<<<grid, Z_NTHREADS, Z_SMEM, s>>>(tmYh, tmYb, tmXh, tmXb, tmA, tmG, p);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        ++launched;
    }
    // 3. the two gradient contractions over G'
    {
        CUtensorMap tmGk, tmGm, tmX, tmY;
        bool ok = encode_map_2d(&tmGk, F32, ws.G, dp.lda, dp.N, (uint64_t)dp.lda * 4, 32, 128, SW) &&
                  encode_map_2d(&tmGm, F32, ws.G, dp.lda, dp.N, (uint64_t)dp.lda * 4, 32, 32, SWA) &&
                  encode_map_2d(&tmX, F32, ws.Xh, Kq, dp.Mp, (uint64_t)Kq * 4, 32, 32, SWA) &&
                  encode_map_2d(&tmY, F32, ws.Yh, Kq, dp.Np, (uint64_t)Kq * 4, 32, 32, SWA);
        if (!ok) return cudaErrorUnknown;
        GradParams g;
        g.Kp = dp.Kp; g.Kc = Kc; g.nkb = (Kc + 31) / 32; g.stop_flag = dp.stop_flag;
        cudaError_t e = cudaFuncSetAttribute(grad_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(grad_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM);
        if (e != cudaSuccess) return e;
        auto grid_of = [&](const GradParams& q) {
            const long long units = (long long)q.n_groups * q.n_steps;
            return units < n_sms ? (int)units : n_sms;
        };
        // dY[j,:] += sum_i G'[j,i] Xh[i,:]
        g.out = dp.dY; g.rows_pad = dp.Np; g.n_groups = (dp.N + 255) / 256; g.n_steps = (dp.M + 31) / 32;
        grad_gemm_kernel<false><<<grid_of(g), G_NTHREADS, G_SMEM, s>>>(tmGk, tmX, g);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        // dX[i,:] += sum_j G'[j,i] Yh[j,:]
        g.out = dp.dX; g.rows_pad = dp.Mp; g.n_groups = (dp.M + 255) / 256; g.n_steps = (dp.N + 31) / 32;
        grad_gemm_kernel<true><<<grid_of(g), G_NTHREADS, G_SMEM, s>>>(tmGm, tmY, g);
        launched += 2;
    }
    if (n_launches) *n_launches = launched;
    return cudaGetLastError();
}

}  // namespace pmf
