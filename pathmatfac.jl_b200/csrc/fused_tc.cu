// Fused data pass on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), K <= 64 (operands zero padded to 64).
//
// One persistent CTA per SM walks a contiguous, cost-balanced range of 128-feature x 64-sample tiles of A
// (ItemIter); the maximal runs inside one feature tile are the ITEMS for which the Y operands and the dY
// accumulator stay resident.  Per tile (32 KB of A, streamed once through a TMA-fed shared-memory ring):
//
//   MMA1  Z[j,i]   = sum_k Y[j,k] X[i,k]     A = Y in TMEM, B = X tile in smem (K-major).  Two-term BF16 split of both
//                                            operands (v = h + l, h = bf16(v), l = bf16(v - h)): Yh*Xh + Yl*Xh + Yh*Xl
//                                            as ONE BF16 contraction over K' = 192 (12 instructions of K = 16; the
//                                            dropped terms are 2^-17 of a product); FP32 accumulate in TMEM
//   epilogue (2 groups of 8 warps on alternating tiles; TMEM lane = feature j, so every column parameter is
//            a per-thread register and every column-gradient sum a private accumulator):
//            ColScale/ColShift, noise loss, dL/dz with the NaN mask (src/layers.jl:9-90, SURVEY.md
//            Appendix B).  dL/dz goes back to TMEM in place of Z and, with 128-bit stores, into shared
//            memory IN PLACE of the A values the same thread just read.  BATCH instantiation: BatchScale /
//            BatchShift too (src/layers.jl:95-214) -- the batch parameters are per-thread registers that change
//            between 16-sample chunks, over sample orders / passes planned on the host (TcBatchDev).
//   MMA2  dX[i,k]  = sum_j G[j,i] Ys[j,k]    A = G in smem read MN-major (M = 64 samples), B = Ys = w_j sigma_j Y
//                                            in smem read MN-major -- no transposed copies exist
//   MMA3  dY[j,k] += sum_i G[j,i] X[i,k]     A = G in TMEM, B = a second copy of the Xh tile read MN-major
//
// Z and dL/dZ never exist in HBM.  dY stays in TMEM across the sample loop of an item; each dX tile is staged
// in shared memory by four drain warps and leaves as ONE TMA reduce-add (no per-lane REDs), two of them in
// flight.  The gradient contractions use rna_tf32 operands (single-pass TF32); precision mode 2 keeps only the
// leading BF16 term of Z (experiments).
//
// Warps: 0-15 epilogue | 16 TMA A | 17 MMA issuer (one elected thread) | 18 TMA X operands | 19 dX reduce-add
// issuer | 20-23 dX drain (TMEM -> staging buffer), one per TMEM lane quarter.  24 warps start with 80 registers;
// setmaxnreg then gives the epilogue warpgroups REGS_EPI and leaves the two small warpgroups REGS_SMALL.
// Shared memory (every box is one 128-byte swizzle row wide):
//   XK  3 stages x 16 KB   Xb tile: bf16 [Xh | Xl] of 64 samples, K-major operand of MMA1, 16-byte-atom swizzle
//   XM  16 KB   rna_tf32(X) tile, 32-byte-atom swizzle: MN-major operand of MMA3
//   YS  32 KB   Ys tile [128 features][64 k], 32-byte-atom swizzle, written by the epilogue warps per item
//   AG  3 stages x 32 KB   A tile as 2 boxes (32 samples x 128 feature rows), 32-byte-atom swizzle; becomes G in place
//   DXS 2 x 16 KB   dX tiles staged for the TMA reduce-add (16-byte-atom swizzle)
// MN-major FP32/TF32 operands exist only in the 32-byte-atom 128B swizzle (UMMA layout type 1).
// TMEM (512 columns): [Yh|Yl] bf16 0 | Z0..Z3 64 | dY 320 | dX0 384 | dX1 448.
// dX accumulators are M = 64 tiles: sample row r lives in lane (r % 16) + 32 * (r / 16).
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

constexpr int BJ = 128, BI = 64, KK = 64;
constexpr int NEPI = 16;                  // epilogue warps: 2 groups x (TMEM lane quarter, 32-column half)
// Warp roles.  The epilogue warps come FIRST: the SM's issue arbiter favours the highest warp id on a
// scheduler, so the four single-thread role warps sit at the top and are never starved by the 16 epilogue
// warps they share schedulers with (a starved MMA thread idles the tensor pipe).
constexpr int W_TMA_A = NEPI, W_MMA = NEPI + 1, W_TMA_X = NEPI + 2, W_MMA1 = NEPI + 3;
// Four more warps, one per TMEM lane quarter, do nothing but move finished dX tiles from TMEM into the staging
// buffers of the TMA reduce-add: off the epilogue warps, whose per-tile cycle sets the pace of the kernel.
constexpr int W_DRAIN0 = NEPI + 4, NDRAIN = 4;
constexpr int NPRE = 4;
constexpr int NTHREADS = 32 * (NPRE + NEPI + NDRAIN);
// 24 warps are launched with 80 registers each (65536 / 768 rounded down to the allocation unit); setmaxnreg moves
// registers inside that allocation: 512 x REGS_EPI + 256 x REGS_SMALL <= 768 x 80.
constexpr int REGS_EPI = 88, REGS_SMALL = 64;
// Pipeline depths.  The X operands come from L2 and are re-loaded while the tensor pipe works on the other
// contractions of the neighbouring tiles, so one stage each suffices; the A/G ring is the HBM stream and
// stays occupied from the TMA issue until MMA2 has consumed G0, so it gets every byte that is left.
#ifndef PMF_SA
#define PMF_SA 3
#endif
constexpr int SA = PMF_SA;
constexpr int SXK = SA == 4 ? 2 : 3, SXM = 1;
constexpr int SZ = SA == 4 ? 3 : 4;       // Z / G0 accumulators in TMEM
constexpr int LA = SZ - 1;                // MMA1 runs LA tiles ahead of MMA2 / MMA3
// The merged X producer loads XK(t) before XM(t - LA), and MMA1 never runs ahead across an item boundary: the last
// MMA3 of an item needs XM(last), which is issued after XK(last + LA), which needs the stage MMA1(last + LA - SXK) frees.
static_assert(SXK >= LA, "with fewer XK stages than the MMA1 look-ahead the X producer deadlocks at item boundaries");
constexpr int SDX = SA == 4 ? 1 : 2;      // dX staging buffers (TMA reduce-adds in flight)
constexpr uint32_t XK_BYTES = 16384 /* bf16 [Xh | Xl] */, XM_BYTES = 16384, YS_BYTES = 32768, AG_BYTES = 32768,
                   DXS_BYTES = 16384 /* dX tile staged for the TMA reduce-add */;
constexpr uint32_t SMEM_DATA = SXK * XK_BYTES + SXM * XM_BYTES + YS_BYTES + SA * AG_BYTES + SDX * DXS_BYTES;
constexpr uint32_t SMEM_TOTAL = SMEM_DATA + 1024 /*align slack*/ + 512 /*barriers*/;
static_assert(SMEM_TOTAL <= 232448, "shared memory budget of one CTA");
constexpr uint32_t TM_YB = 0, TM_Z0 = 64, TM_DY = 320, TM_DX0 = 384;   // [Yh|Yl] bf16 pairs; Z_b = Z0+64b, dX1 = dX0+64
static_assert(TM_Z0 + 64 * SZ <= TM_DY, "TMEM map");

enum Bar { B_FULL_XK = 0, B_EMPTY_XK = B_FULL_XK + SXK, B_FULL_XM = B_EMPTY_XK + SXK, B_EMPTY_XM = B_FULL_XM + SXM,
           B_FULL_A = B_EMPTY_XM + SXM, B_EMPTY_AG = B_FULL_A + SA, B_Z_FULL = B_EMPTY_AG + SA, B_G_READY = B_Z_FULL + SZ,
           B_DX_FULL = B_G_READY + SZ, B_DX_EMPTY = B_DX_FULL + 2, B_Y_READY = B_DX_EMPTY + 2, B_DY_FULL, B_DY_EMPTY,
           B_DXS_FULL, B_DXS_DONE = B_DXS_FULL + SDX, B_Z_EMPTY = B_DXS_DONE + SDX, B_COUNT = B_Z_EMPTY + SZ };
static_assert(8 * B_COUNT + 8 <= 512, "barrier area");

using namespace tcx;

// ordinal / hinge noise models: rare on the hot path, kept out of line to bound code size.
// Returns (loss, dloss/dz); (0, 0) for a missing entry.
__device__ __noinline__ float2 noise_eval_slow(int dist, float z, float a, const float* __restrict__ th_range,
                                                float ord_eps, float margin) {
    if (!is_observed(a)) return make_float2(0.f, 0.f);
    const float4 th4 = __ldg(reinterpret_cast<const float4*>(th_range));
    const float th[4] = {th4.x, th4.y, th4.z, th4.w};
    float l, g;
    noise_eval(dist, z, a, th, ord_eps, margin, l, g);
    return make_float2(l, g);
}

__device__ __noinline__ float2 threshold_grads_slow(int dist, float z, float a, const float* __restrict__ th_range,
                                                    float ord_eps, float margin) {
    if (!is_observed(a)) return make_float2(0.f, 0.f);
    const float4 th4 = __ldg(reinterpret_cast<const float4*>(th_range));
    const float th[4] = {th4.x, th4.y, th4.z, th4.w};
    float g1, g2;
    noise_threshold_grads(dist, z, a, th, ord_eps, margin, g1, g2);
    return make_float2(g1, g2);
}

struct TcParams {
    DataPassParams dp;
    int n_jt, n_it;
    int z_passes;      // 3 = Yh Xh + Yl Xh + Yh Xl (split-BF16 Z at FP32 level), 1 = the leading term Yh Xh only
    int ablate;        // PMF_TC_ABLATE (performance experiments only; results are wrong when non-zero)
    long long* trace;  // PMF_TC_TRACE: per-tile clock64 stamps of one CTA (16 events x TRACE_TILES), else null
    int trace_cta;
    int flags;         // PMF_TC_FLAGS: trace filters (16 = only the per-tile G_READY stamp, 32 = only the MMA thread)
    // batch layers (BATCH instantiation): n_jt counts PASSES = (feature tile, sample order) pairs, see TcBatchDev
    const int32_t* pass_feat0;   // [n_jt] first feature of the pass' tile
    const int32_t* pass_order;   // [n_jt] sample order of the pass
    const int32_t* view_order;   // [n_views] sample order in which the view's batches are contiguous
    const uint16_t* boc;         // [n_views][Mp/16] batch of the 16 positions of a chunk, in the view's own order
    int n_views, n_orders;
};
constexpr uint32_t B_NONE = 0xFFFFu;
constexpr int TRACE_TILES = 96, TRACE_EV = 32, TRACE_CTAS = 160;   // + per-CTA (start, end) clocks after the stamps

// Work distribution.  The tiles, flattened feature-tile-major ([jt][it]), are cut into gridDim.x equal
// contiguous ranges, one per CTA; a range is walked as ITEMS = maximal runs inside one feature tile (the Y
// operands and the dY accumulator stay resident for an item).  Every CTA gets the same number of tiles
// (+-1) and crosses at most a handful of feature-tile boundaries, where the pipeline has to drain.
// The cut points are equidistant in COST: a tile of bernoulli columns (three MUFU per entry) takes longer
// than a tile of normal columns, and a CTA's whole range may lie in one assay (tc_cost_cum, built by
// pmf_set_noise, holds the cumulated per-feature-tile cost).
struct ItemIter {
    int t, t_end, n_it;
    int jt, it0, it1;
    // tile index at cost position  frac = num / den  of the whole pass.  A feature tile costs its tiles plus the
    // pipeline drain / refill at its boundary (CB, in the units of tc_cost_cum: a normal tile = 100).
    static constexpr long long CB = 500;
    static __device__ __forceinline__ int cut(const TcParams& p, unsigned num, unsigned den) {
        const int32_t* cum = p.dp.tc_cost_cum;
        if (num >= den) return p.n_jt * p.n_it;
        auto upto = [&](int jt) { return (long long)cum[jt] * p.n_it + CB * jt; };   // cost of the feature tiles before jt
        const long long x = upto(p.n_jt) * num / den;
        int lo = 0, hi = p.n_jt;                                    // feature tile with upto(jt) <= x < upto(jt + 1)
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (upto(mid) <= x) lo = mid; else hi = mid;
        }
        const long long w = cum[lo + 1] - cum[lo];
        long long it = (x - upto(lo) - CB) / (w > 0 ? w : 1);
        if (it < 0) it = 0;
        if (it > p.n_it) it = p.n_it;
        return lo * p.n_it + (int)it;
    }
    __device__ __forceinline__ explicit ItemIter(const TcParams& p) {
        t = cut(p, blockIdx.x, gridDim.x);
        t_end = cut(p, blockIdx.x + 1, gridDim.x);
        n_it = p.n_it;
        jt = it0 = it1 = 0;
    }
    __device__ __forceinline__ bool next() {
        if (t >= t_end) return false;
        jt = t / n_it;
        it0 = t - jt * n_it;
        const int len = min(n_it - it0, t_end - t);
        it1 = it0 + len;
        t += len;
        return true;
    }
    __device__ __forceinline__ int peek_jt() const { return t < t_end ? t / n_it : -1; }   // feature tile of the NEXT item
};

// stage / parity bookkeeping of a ring of `n` buffers (avoids a modulo per tile)
struct Ring {
    uint32_t s = 0, ph = 0;
    __device__ __forceinline__ void next(uint32_t n) { if (++s == n) { s = 0; ph ^= 1u; } }
};

// THR: also accumulate the gradients of the interior ordinal thresholds (update_noise_models on a model with ordinal
// columns).  A separate instantiation: the production kernel carries none of that code.
template <bool DBG, bool BATCH, bool THR>
__global__ void __launch_bounds__(NTHREADS, 1)
data_pass_tc_kernel(const __grid_constant__ CUtensorMap tmXb, const __grid_constant__ CUtensorMap tmXm,
                    const __grid_constant__ CUtensorMap tmA,
                    const __grid_constant__ CUtensorMap tmDX, const TcParams p) {
    const DataPassParams& dp = p.dp;
    if (dp.stop_flag != nullptr && *dp.stop_flag != 0) return;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t XK = base, XM = XK + SXK * XK_BYTES, YS = XM + SXM * XM_BYTES, AG = YS + YS_BYTES, DXS = AG + SA * AG_BYTES;
    const uint32_t BARS = base + SMEM_DATA;
    uint8_t* ys_ptr = gbase + (YS - base);
    uint8_t* ag_ptr0 = gbase + (AG - base);
    uint8_t* dxs_ptr = gbase + (DXS - base);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + SMEM_DATA + 8 * B_COUNT);
    __shared__ double red_smem[NTHREADS / 32];
    auto bar = [&](int b) { return BARS + 8u * (uint32_t)b; };
    // event stamp of CTA 0 (timeline experiments; the branch is CTA-uniform)
    auto stamp = [&](uint32_t gg, int ev) {
        if (DBG && p.trace != nullptr && blockIdx.x == (unsigned)p.trace_cta && gg < (uint32_t)TRACE_TILES &&
            (!(p.flags & 16) || ev == 9) && (!(p.flags & 32) || ev == 1 || ev == 2 || ev == 3 || ev == 4 || ev >= 13) &&
            (!(p.flags & 64) || (ev >= 5 && ev <= 9) || (ev >= 16 && ev <= 21)) &&
            (!(p.flags & 128) || ev == 0 || (ev >= 5 && ev <= 9)))
            p.trace[gg * TRACE_EV + ev] = clock64();
    };

    // warp index through a shuffle: the compiler then treats role branches as warp-uniform and keeps
    // MMA descriptors / barrier addresses in uniform registers
    if (p.trace != nullptr && threadIdx.x == 0 && blockIdx.x < (unsigned)TRACE_CTAS)     // CTA-uniform; also in production (PMF_TC_CTATIMES)
        p.trace[TRACE_TILES * TRACE_EV + 2 * blockIdx.x] = (long long)globaltimer_ns();
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int b = 0; b < B_COUNT; ++b) {
            uint32_t cnt = 1u;
            // Epilogue warps arrive ONCE PER WARP (lane 0, after the lanes' fences and a __syncwarp): an arrive is
            // an atomic on the barrier word, and 256 of them per barrier and tile serialise in the shared-memory unit
            if (b == B_Y_READY || b == B_DY_EMPTY) cnt = NEPI;                             // every epilogue warp
            if (b >= B_G_READY && b < B_G_READY + SZ) cnt = NEPI / 2;                      // one epilogue group
            if (b == B_DX_EMPTY || b == B_DX_EMPTY + 1 || (b >= B_DXS_FULL && b < B_DXS_FULL + SDX)) cnt = NDRAIN;   // the drain warps
            mbar_init(bar(b), cnt);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = *tmem_slot;

    // Register re-allocation: ONE setmaxnreg for the two small warpgroups (warps 16-23) and one for the four epilogue
    // warpgroups -- every warp of a warpgroup executes the same instruction -- each dominating the code of its roles.
    if (warp >= NEPI) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_SMALL));
    if (warp < NEPI) {
        // (epilogue: the last branch)
    } else if (warp == W_TMA_A || warp == W_TMA_X) {
        // ================================ TMA producers ===========================================
        // Warp W_TMA_A streams the A tiles (HBM); warp W_TMA_X feeds both X operand buffers (L2): the K-major pair
        // (Xh | Xl) of tile t, then the MN-major Xh copy of tile t-2 -- the order in which MMA1 (two tiles
        // ahead) and MMA3 consume them, so neither load waits behind a buffer that is released later.
        if (lane == 0 && warp == W_TMA_A) {
            Ring r;
            uint32_t gcount = 0;
            for (ItemIter itx(p); itx.next();) {
                const int it0 = itx.it0, it1 = itx.it1;
                const int j0 = itx.jt * BJ;
                for (int it = it0; it < it1; ++it, r.next(SA), ++gcount) {
                    const int i0 = it * BI;
                    mbar_wait(bar(B_EMPTY_AG + r.s), r.ph ^ 1);
                    stamp(gcount, 0);
                    mbar_expect_tx(bar(B_FULL_A + r.s), AG_BYTES);
                    for (int iq = 0; iq < 2; ++iq)
                        tma_load_2d(AG + r.s * AG_BYTES + iq * 16384, &tmA, bar(B_FULL_A + r.s),
                                    (DBG && (p.ablate & 32)) ? 32 * iq : i0 + 32 * iq, (DBG && (p.ablate & 32)) ? 0 : j0);
                    // The load of a tile can only be issued when its ring stage is free, about 1.5 tile periods before
                    // the epilogue wants it: the HBM latency would sit inside the stage's cycle.  An L2 prefetch a few
                    // tiles further ahead costs no stage and turns that latency into an L2 hit.
                    const int pf = DBG ? (p.flags >> 8) & 15 : 0;
                    if (pf && it + pf < it1) {
                        tma_prefetch_2d(&tmA, i0 + 64 * pf, j0);
                        tma_prefetch_2d(&tmA, i0 + 64 * pf + 32, j0);
                    }
                }
            }
        } else if (lane == 0) {
            Ring rk, rm;
            int pend[LA];                 // sample offsets of the last LA tiles whose MN-major copy is still due
#pragma unroll
            for (int k = 0; k < LA; ++k) pend[k] = -1;
            auto load_xm = [&](int i0) {
                mbar_wait(bar(B_EMPTY_XM + rm.s), rm.ph ^ 1);
                if (DBG && (p.ablate & 128)) {       // experiment: no L2 -> shared-memory traffic for the MN-major copy
                    mbar_arrive(bar(B_FULL_XM + rm.s));
                } else {
                    mbar_expect_tx(bar(B_FULL_XM + rm.s), XM_BYTES);
                    for (int kb = 0; kb < 2; ++kb)
                        tma_load_2d(XM + rm.s * XM_BYTES + kb * 8192, &tmXm, bar(B_FULL_XM + rm.s), 32 * kb, i0);
                }
                rm.next(SXM);
            };
            for (ItemIter itx(p); itx.next();) {
                const int it0 = itx.it0, it1 = itx.it1;
                const int xrow0 = BATCH ? __ldg(p.pass_order + itx.jt) * dp.Mp : 0;   // this order's copy of the X operands
                for (int it = it0; it < it1; ++it) {
                    const int i0 = xrow0 + it * BI;
                    mbar_wait(bar(B_EMPTY_XK + rk.s), rk.ph ^ 1);
                    if (DBG && (p.ablate & 64)) {    // experiment: no L2 -> shared-memory traffic for the K-major pair
                        mbar_arrive(bar(B_FULL_XK + rk.s));
                    } else {
                        mbar_expect_tx(bar(B_FULL_XK + rk.s), XK_BYTES);
                        const uint32_t dst = XK + rk.s * XK_BYTES;
                        for (int kb = 0; kb < 2; ++kb)
                            tma_load_2d(dst + kb * 8192, &tmXb, bar(B_FULL_XK + rk.s), 64 * kb, i0);   // bf16 Xh | Xl
                    }
                    rk.next(SXK);
                    if (pend[0] >= 0) load_xm(pend[0]);
#pragma unroll
                    for (int k = 0; k + 1 < LA; ++k) pend[k] = pend[k + 1];
                    pend[LA - 1] = i0;
                }
            }
#pragma unroll
            for (int k = 0; k < LA; ++k)
                if (pend[k] >= 0) load_xm(pend[k]);
        }
    } else if (warp == W_MMA1) {
      if (elect_one()) {
        // ================================ MMA issuer 1: Z = Y X' =====================================
        // tcgen05.mma is accepted at the rate the tensor pipe executes it (the queue holds one or two instructions),
        // so every cycle an issuing thread spends in a barrier wait is a cycle the pipe idles.  The contractions
        // are therefore issued by TWO threads: this one runs MMA1 up to SZ tiles ahead, the other one MMA2 / MMA3;
        // while one of them waits, the instructions of the other keep the pipe busy.  The order the single
        // in-order thread used to provide -- MMA3 of tile t - SZ has read G out of the Z buffer that MMA1 of tile
        // t overwrites -- is now the barrier Z_EMPTY.
        // BF16 x BF16 -> F32 (kind::f16): c_format F32, a_format = b_format = BF16, both K-major
        const uint32_t id_zb = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const int kn = dp.Kp;                                       // multiple of 8, <= 64
        if (tm != 0u) __trap();
        const uint32_t tmu = 0u;
        auto kstep = [](uint64_t d0, int s, uint32_t atom) { return d0 + (uint64_t)((((s >> 2) * atom) + (s & 3) * 32) >> 4); };
        uint32_t q = 0, g1 = 0;
        Ring rx1, rz1;         // XK stage / Z buffer of the next MMA1
        // This thread runs tiles ahead of everything else, so it also issues the TMA reduce-adds of the staged dX
        // tiles, RLAG tiles behind its own MMA1 (by then the drain warps have staged that tile): two bulk groups in
        // flight, the staging buffer of a tile is handed back once the engine has read it.  Pending reduce-adds
        // are flushed at the end of an item -- the drain of the item's last tiles needs the buffers back before
        // the item can complete, and this thread is about to block on the next item's Y operands.
        constexpr int RLAG = SDX == 2 ? 5 : 3;
        uint32_t gr = 0;       // next tile whose dX is to be reduced
        auto reduce_tile = [&](int it_r, int xrow0) {
            const uint32_t sb = gr % SDX;
            mbar_wait(bar(B_DXS_FULL + sb), (gr / SDX) & 1);
            stamp(gr, 31);
            if (!(DBG && (p.ablate & 16))) {
                tma_reduce_add_2d(&tmDX, DXS + sb * DXS_BYTES, 0, xrow0 + it_r * BI);
                if (dp.Kp > 32) tma_reduce_add_2d(&tmDX, DXS + sb * DXS_BYTES + 8192, 32, xrow0 + it_r * BI);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (SDX == 2) {
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                if (gr > 0) mbar_arrive(bar(B_DXS_DONE + (sb ^ 1u)));     // tile gr - 1 has been read
            } else {
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                mbar_arrive(bar(B_DXS_DONE));
            }
            stamp(gr, 12);
            ++gr;
        };
        for (ItemIter itx(p); itx.next();) {
            const int it0 = itx.it0, it1 = itx.it1;
            const int xrow0 = BATCH ? __ldg(p.pass_order + itx.jt) * dp.Mp : 0;   // this order's copy of dX
            mbar_wait(bar(B_Y_READY), q & 1);
            int it_r = it0;
            for (int it = it0; it < it1; ++it, ++g1) {
                stamp(g1, 13);
                mbar_wait(bar(B_Z_EMPTY + rz1.s), rz1.ph ^ 1);
                mbar_wait(bar(B_FULL_XK + rx1.s), rx1.ph);
                tc_fence_after();
                stamp(g1, 1);
                const uint32_t zt = tmu + TM_Z0 + 64 * rz1.s;
                // Xb tile: box 0 = Xh (64 bf16 = 128 bytes per sample row), box 1 = Xl; a k-step of 16 is 32 bytes
                const uint64_t xb = umma_desc_k(XK + rx1.s * XK_BYTES);
#pragma unroll
                for (int s = 0; s < 4; ++s)       // Yh * Xh
                    if (16 * s < kn) mma_ts_f16(zt, tmu + TM_YB + 8 * s, kstep(xb, s, 8192), id_zb, s > 0 ? 1u : 0u);
                if (p.z_passes == 3) {
#pragma unroll
                    for (int s = 0; s < 4; ++s)   // Yl * Xh
                        if (16 * s < kn) mma_ts_f16(zt, tmu + TM_YB + 32 + 8 * s, kstep(xb, s, 8192), id_zb, 1u);
#pragma unroll
                    for (int s = 0; s < 4; ++s)   // Yh * Xl
                        if (16 * s < kn) mma_ts_f16(zt, tmu + TM_YB + 8 * s, kstep(xb, 4 + s, 8192), id_zb, 1u);
                }
                tc_commit_elect(bar(B_EMPTY_XK + rx1.s));
                tc_commit_elect(bar(B_Z_FULL + rz1.s));
                stamp(g1, 14);
                rx1.next(SXK);
                rz1.next(SZ);
                if (it - it_r >= RLAG) reduce_tile(it_r++, xrow0);
            }
            while (it_r < it1) reduce_tile(it_r++, xrow0);
            ++q;
        }
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (SDX == 2 && gr > 0) mbar_arrive(bar(B_DXS_DONE + ((gr - 1u) & 1u)));
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // every reduce-add has been performed
      }
      __syncwarp();
    } else if (warp == W_MMA) {
      if (elect_one()) {
        // ================================ MMA issuer 2: dX, dY ======================================
        // K < 64: the gradient tiles are only Kp columns wide
        const int kn = dp.Kp;                                       // multiple of 8, <= 64
        const uint32_t id_dx = umma_idesc(64, kn, true, true);     // dX = G0'(smem, MN) * Yh(smem, MN)
        const uint32_t id_dy = umma_idesc(128, kn, false, true);   // dY = G0(tmem) * Xh(smem, MN)
        // All 512 columns are allocated by the only CTA on the SM, so the TMEM base is 0 (checked): a literal
        // keeps every tcgen05.mma operand in uniform registers (no per-instruction R2UR).
        if (tm != 0u) __trap();
        const uint32_t tmu = 0u;
        uint32_t g = 0, q = 0;
        Ring rx3, ra, rz;      // XM stage of the next MMA3, A/G stage of the next MMA2, Z buffer of the next MMA2/3
        for (ItemIter itx(p); itx.next();) {
            const int it0 = itx.it0, it1 = itx.it1;
            mbar_wait(bar(B_Y_READY), q & 1);
            mbar_wait(bar(B_DY_EMPTY), (q & 1) ^ 1);
            for (int it = it0; it < it1; ++it, ++g) {
                const uint32_t b = g & 1, ph = (g >> 1) & 1;
                mbar_wait(bar(B_G_READY + rz.s), rz.ph);
                stamp(g, 2);
                mbar_wait(bar(B_DX_EMPTY + b), ph ^ 1);
                tc_fence_after();
                stamp(g, 3);
                {
                    // MMA2 first (dX = G0' * Yh): its completion releases the A/G buffer for the TMA producer
                    const uint64_t gd = umma_desc_mn(AG + ra.s * AG_BYTES, 16384u), yd = umma_desc_mn(YS, 16384u);
                    const uint32_t dxt = tmu + TM_DX0 + 64 * b;
#pragma unroll
                    for (int s = 0; s < 16; ++s) {
                        if (DBG && (p.ablate & 1) && s > 0) break;
                        mma_ss(dxt, gd + (uint64_t)(s * 64), yd + (uint64_t)(s * 64), id_dx, s > 0 ? 1u : 0u);
                    }
                    tc_commit_elect(bar(B_EMPTY_AG + ra.s));
                    tc_commit_elect(bar(B_DX_FULL + b));
                    ra.next(SA);
                }
                {
                    // MMA3: dY += G0 * Xh  (MN-major copy of the Xh tile)
                    mbar_wait(bar(B_FULL_XM + rx3.s), rx3.ph);
                    tc_fence_after();
                    stamp(g, 4);
                    const uint32_t ga = tmu + TM_Z0 + 64 * rz.s;
                    const uint64_t xd = umma_desc_mn(XM + rx3.s * XM_BYTES, 8192u);
                    const uint32_t first = it > it0 ? 1u : 0u;
#pragma unroll
                    for (int s = 0; s < 8; ++s) {
                        if (DBG && (p.ablate & 2) && s > 0) break;
                        mma_ts(tmu + TM_DY, ga + 8 * s, xd + (uint64_t)(s * 64), id_dy, s > 0 ? 1u : first);
                    }
                    tc_commit_elect(bar(B_EMPTY_XM + rx3.s));
                    tc_commit_elect(bar(B_Z_EMPTY + rz.s));
                    stamp(g, 15);
                    rx3.next(SXM);
                    rz.next(SZ);
                }
            }
            tc_commit_elect(bar(B_DY_FULL));
            ++q;
        }
      }
      __syncwarp();
    } else if (warp >= W_DRAIN0) {
        // ================================ dX drain warps ===========================================
        // dX tile of EVERY tile: TMEM -> registers -> one of the two swizzled staging buffers, from where the MMA1
        // thread issues a TMA reduce-add into global dX (the L2 does the additions on full lines; no per-lane REDs
        // through the LSU), two of them in flight.  M = 64 accumulator: sample row r sits in lane (r % 16) +
        // 32 * (r / 16), so warp q of the four holds rows 16q .. 16q + 15 in the first 16 lanes of its quarter:
        // ONE tcgen05.ld of shape 16x256b.x8 brings the warp's 16 x 64 block in, spread over all 32 threads
        // (thread t: rows t/4 and t/4 + 8, columns 8j + 2(t%4), +1 of every 8-column atom j).
        const int quarter = warp & 3;
        const uint32_t lane_addr = ((uint32_t)(32 * quarter)) << 16;
        const int ra_ = 16 * quarter + (lane >> 2), rb_ = ra_ + 8;   // sample rows of the tile held by this thread
        const uint32_t cw = (uint32_t)(lane & 3) * 8u;                // byte offset of its column pair inside a 32-byte group
        const bool tr = warp == W_DRAIN0 && lane == 0;
        uint32_t g = 0;
        for (ItemIter itx(p); itx.next();) {
            const int it0 = itx.it0, it1 = itx.it1;
            for (int it = it0; it < it1; ++it, ++g) {
                const uint32_t b = g & 1;                  // dX accumulator
                const uint32_t sb = g % SDX;               // staging buffer
                uint8_t* const buf = dxs_ptr + sb * DXS_BYTES;
                // experiment (PMF_TC_FLAGS bits 8..11 = distance in tiles): L2 prefetch of a later A tile through the LSU
                if (!BATCH) {
                    const int pf = (p.flags >> 8) & 15;
                    if (pf && it + pf < it1) {
                        const int jrow = itx.jt * BJ + 32 * quarter + lane;
                        if (jrow < dp.N) {
                            const float* pa = dp.A + (size_t)jrow * dp.lda + (size_t)(it + pf) * BI;
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(pa));
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + 32));
                        }
                    }
                }
                if (tr) stamp(g, 10);
                mbar_wait(bar(B_DX_FULL + b), (g >> 1) & 1);
                tc_fence_after();
                if (tr) stamp(g, 11);
                uint32_t r0[32];
                TMEM_LD_16x256b_x8(tm + lane_addr + TM_DX0 + 64 * b, r0);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                // the accumulator has been read: MMA2 of tile g + 2 may overwrite it
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_relaxed(bar(B_DX_EMPTY + b));
                if (tr) stamp(g, 28);
                // the reduce-add of tile g - 2 has read this staging buffer (first use: parity 1 passes on a fresh barrier)
                mbar_wait(bar(B_DXS_DONE + sb), ((g / SDX) - 1u) & 1u);
                if (tr) stamp(g, 29);
                if (!(DBG && (p.ablate & 16))) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        // columns 8j + 2(t%4), +1: box j / 4, 16-byte chunk 2 (j % 4) + (t % 4) / 2 of the 128-byte row
                        const uint32_t c = 2u * (j & 3) + ((lane & 3) >> 1);
                        uint8_t* const bx = buf + (j >> 2) * 8192;
                        *reinterpret_cast<uint2*>(bx + ra_ * 128 + ((c ^ (ra_ & 7)) << 4) + (cw & 8u)) = make_uint2(r0[4 * j], r0[4 * j + 1]);
                        *reinterpret_cast<uint2*>(bx + rb_ * 128 + ((c ^ (rb_ & 7)) << 4) + (cw & 8u)) = make_uint2(r0[4 * j + 2], r0[4 * j + 3]);
                    }
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(B_DXS_FULL + sb));
                if (tr) stamp(g, 30);
            }
        }
    }
    if (warp < NEPI) {
        // ================================ epilogue warps ===========================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_EPI));
        // Two groups of 8 warps work on alternating tiles, so that the load / compute / store-drain
        // phases of one tile overlap those of the next.  Inside a tile a warp owns (TMEM lane quarter,
        // 32-sample half): TMEM lanes 32*quarter.., columns 32*h32.. = one row of A box h32 per thread.
        // Item prologue / flush (Y operands in, dY tile out) are spread over all 16 warps by 16-column chunk.
        const int quarter = warp & 3;
        const int c16 = warp >> 2;                    // 0..3
        const int grp = c16 >> 1, h32 = c16 & 1;
        const int lrow = 32 * quarter + lane;         // feature row of the tile
        const uint32_t lane_addr = ((uint32_t)(32 * quarter)) << 16;
        // byte offset of logical 16-byte chunk c of this thread's 128-byte box row:
        // 32-byte slot (c>>1) ^ (row&3), half c&1   (128B swizzle with 32-byte atoms)
        auto chunk_off = [&](int c) { return (uint32_t)(lrow * 128 + ((((c >> 1) ^ (lane & 3)) << 5) | ((c & 1) << 4))); };
        const uint32_t tile_box = (uint32_t)h32 * 16384u;
        const bool swapped = (lane & 4) != 0;
        const uint32_t half_swap = swapped ? 16u : 0u;
        uint32_t g = 0, q = 0;
        Ring ra, rz;
        if (grp == 1) { ra.next(SA); rz.next(SZ); }
        double loss_d = 0.0;

        // Per-item operands of this thread (lane = feature): column constants and its 16-column chunk of
        // the Y row.  They are fetched one item AHEAD, right before the wait for the current item's last
        // contraction, so the two global round trips are off the item-to-item critical path.
        struct ItemRegs { float logsigma, mu, w; int ci; int boff, bview; float4 y[4]; };
        auto load_item = [&](int jt_, ItemRegs& r) {
            const int feat0 = BATCH ? __ldg(p.pass_feat0 + jt_) : jt_ * BJ;
            const int j_ = feat0 + lrow;
            const int jj_ = j_ < dp.N ? j_ : dp.N - 1;   // padding rows follow the last column (same noise model: no divergence)
            r.logsigma = __ldg(dp.logsigma + jj_);
            r.mu = __ldg(dp.mu + jj_);
            r.w = j_ < dp.N ? __ldg(dp.weight + jj_) : 0.f;
            r.ci = __ldg(dp.colinfo + jj_);
            if (BATCH) {
                // offset of the column's block of the batch tables (-1: the column's view has no batch layers)
                // and the view whose sample tables the column reads.  A column whose view needs another sample
                // order is all-NaN in this pass (it is served by the tile's other pass): no batch bookkeeping.
                int bo = j_ < dp.N ? __ldg(dp.bcol_off + jj_) : -1;
                const int bv = bo >= 0 ? __ldg(dp.bcol_view + jj_) : 0;
                if (bo >= 0 && __ldg(p.view_order + bv) != __ldg(p.pass_order + jt_)) bo = -1;
                r.boff = bo;
                r.bview = bv;
            }
            // K <= 64: the factor arrays keep their own pitch Kp; columns k >= Kp are zeros here, in the
            // operand scratch (Xh / Xb are 64 / 128 wide and zero padded) and in every TMA box (OOB fill)
            const float4* yrow = reinterpret_cast<const float4*>(dp.Y + (size_t)j_ * dp.Kp) + 4 * c16;
#pragma unroll
            for (int v = 0; v < 4; ++v)
                r.y[v] = (16 * c16 + 4 * v < dp.Kp) ? __ldg(yrow + v) : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        // L2 prefetch of the same addresses, issued a whole item earlier (the A stream evicts them from L2)
        auto prefetch_item = [&](int jt_) {
            const int j_ = (BATCH ? __ldg(p.pass_feat0 + jt_) : jt_ * BJ) + lrow;
            const int jj_ = j_ < dp.N ? j_ : dp.N - 1;   // padding rows follow the last column (same noise model: no divergence)
            if (c16 == 0) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(dp.logsigma + jj_));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(dp.mu + jj_));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(dp.weight + jj_));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(dp.colinfo + jj_));
            }
            if (16 * c16 < dp.Kp)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(dp.Y + (size_t)j_ * dp.Kp + 16 * c16));
        };
        ItemRegs cur;
        {
            const int jt_first = ItemIter(p).peek_jt();
            if (jt_first >= 0) load_item(jt_first, cur);
        }
        // column-side results of the previous item, reduced into global memory one item late (see the item
        // epilogue): 128-bit REDs for the dY tile (several CTAs contribute to a feature tile)
        float4 pend_dy[4];
        float pend_dmu = 0.f, pend_dls = 0.f;
        int pend_j = -1;
        float pend_dth = 0.f, pend_dld = 0.f;    // BATCH: last batch segment of the previous item
        int pend_bidx = -1;
        auto store_partials = [&]() {
            if (pend_j >= 0) {
                float4* dst = reinterpret_cast<float4*>(dp.dY + (size_t)pend_j * dp.Kp + 16 * c16);
#pragma unroll
                for (int v = 0; v < 4; ++v)
                    if (16 * c16 + 4 * v < dp.Kp) atomicAdd(dst + v, pend_dy[v]);
                atomicAdd(dp.dmu + pend_j, pend_dmu);
                atomicAdd(dp.dlogsigma + pend_j, pend_dls);
            }
            pend_j = -1;
        };
        // dtheta / dlogdelta of a closed batch segment: issued right AFTER the tile's G_READY arrive (or the next
        // item's Y_READY arrive), never in front of one
        auto store_segment = [&]() {
            if (BATCH && pend_bidx >= 0) {
                atomicAdd(dp.dtheta + pend_bidx, pend_dth);
                atomicAdd(dp.dlogdelta + pend_bidx, pend_dld);
                pend_bidx = -1;
            }
        };

        for (ItemIter itx(p); itx.next();) {
            const int jt = itx.jt, it0 = itx.it0, it1 = itx.it1;
            const int j = (BATCH ? __ldg(p.pass_feat0 + jt) : jt * BJ) + lrow;
            const bool jok = j < dp.N;
            // per-thread column constants (lane = feature)
            const float sigma = __expf(cur.logsigma);
            const float muj = cur.mu;
            const float wj = cur.w;
            const int ci = cur.ci;
            const int dist = ci & 0xff;
            const float* th_range = dp.thresholds + 4 * (ci >> 8);
            // G0 = (w_j sigma_j) * dloss/dz4 (times delta_bj with batch layers, folded into the entries).  The per-feature factor never
            // touches the per-entry path: MMA2 reads Yh scaled by it, the dY tile is scaled at the flush.
            const float gscale = sigma * wj;
            float dmu_acc = 0.f, loss_acc = 0.f;   // unweighted sums over this item (<= a few thousand terms)
            // Batch layers (src/layers.jl:221-253, src/batch_array.jl:132-212): z4 = z sigma_j delta_bj + mu_j + theta_bj.
            // The 16 samples of a chunk are the same for every thread of the warp and, in the sample order of the
            // column's view, ALWAYS lie in one batch (boc; the orders pad every batch to a multiple of 16
            // positions).  The batch parameters are therefore per-thread registers that change at SEGMENT
            // boundaries only; a segment's sums  sum g  and  sum g z  give dtheta, dlogdelta = sigma delta sum g z,
            // and its share of dmu and dlogsigma (sum g delta).  delta is folded into the dloss/dz entries.
            const int boff = BATCH ? cur.boff : -1;
            const uint16_t* const boc_row = BATCH ? p.boc + (size_t)cur.bview * (dp.Mp >> 4) : nullptr;
            uint32_t cur_b = B_NONE;               // no batch parameters in force
            float sd = sigma, mt = muj, dl = 1.f;  // sigma_j delta_bj, mu_j + theta_bj, delta_bj
            float seg_g = 0.f, seg_gz = 0.f, dls_acc = 0.f;
            uint32_t pre_b = B_NONE;               // batch whose parameters were fetched ahead (per tile)
            float pre_ld = 0.f, pre_th = 0.f;
            auto close_segment = [&]() {
                if (cur_b != B_NONE) {
                    store_segment();               // a still older segment (several switches inside one tile)
                    pend_bidx = boff + (int)cur_b;
                    pend_dth = seg_g * wj;
                    pend_dld = seg_gz * sd * wj;
                }
                dmu_acc += seg_g;
                dls_acc = fmaf(seg_g, dl, dls_acc);
                seg_g = seg_gz = 0.f;
            };
            // batches of this thread's two 16-sample chunks of tile it_ (16 bits each) and the parameters of the first
            // one that differs from the batch in force: fetched a whole tile ahead, off the critical path
            uint32_t ids = B_NONE * 0x10001u;
            auto fetch_ids = [&](int it_) {
                ids = __ldg(reinterpret_cast<const uint32_t*>(boc_row + ((it_ * BI + 32 * h32) >> 4)));
                pre_b = (ids & 0xffffu) != cur_b ? (ids & 0xffffu) : (ids >> 16);
                if (pre_b != cur_b) {
                    pre_ld = __ldg(dp.logdelta + boff + (int)pre_b);
                    pre_th = __ldg(dp.theta + boff + (int)pre_b);
                }
            };
            if (BATCH && boff >= 0) {
                const int it_first = it0 + (int)((g ^ (uint32_t)grp) & 1u);     // this group's first tile of the item
                if (it_first < it1) fetch_ids(it_first);
            }
            auto enter_segment = [&](uint32_t b_) {
                close_segment();
                cur_b = b_;
                const float ld = b_ == pre_b ? pre_ld : __ldg(dp.logdelta + boff + (int)b_);
                const float th = b_ == pre_b ? pre_th : __ldg(dp.theta + boff + (int)b_);
                dl = __expf(ld);
                sd = sigma * dl;
                mt = muj + th;
            };

            // ---- Y tile: h = rna_tf32(y), l = y - h -> TMEM (A of MMA1); gscale * y -> shared memory (B of MMA2).
            // The previous item's MMAs have all completed (B_DY_FULL was waited on), so both are free.
            {
                uint32_t lob[8], hib[8];      // packed bf16 pairs of Yh = bf16(y) and of Yl = bf16(y - Yh)
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const float4 y4 = cur.y[v];
                    const uint32_t h01 = pack_bf16(y4.x, y4.y), h23 = pack_bf16(y4.z, y4.w);
                    hib[2 * v] = h01;
                    hib[2 * v + 1] = h23;
                    lob[2 * v] = pack_bf16(y4.x - __uint_as_float(h01 << 16), y4.y - __uint_as_float(h01 & 0xffff0000u));
                    lob[2 * v + 1] = pack_bf16(y4.z - __uint_as_float(h23 << 16), y4.w - __uint_as_float(h23 & 0xffff0000u));
                    *reinterpret_cast<uint4*>(ys_ptr + (uint32_t)(c16 >> 1) * 16384u + chunk_off(4 * (c16 & 1) + v)) =
                        make_uint4(rna_tf32(gscale * y4.x), rna_tf32(gscale * y4.y), rna_tf32(gscale * y4.z), rna_tf32(gscale * y4.w));
                }
                const bool trp = quarter == 0 && h32 == 0 && lane == 0;
                if (trp && !(p.flags & 64)) stamp(g, 19 + 6 * grp);
                TMEM_ST8(tm + lane_addr + TM_YB + 8 * c16, hib);          // Yh: 16 k of this warp = 8 columns of bf16 pairs
                TMEM_ST8(tm + lane_addr + TM_YB + 32 + 8 * c16, lob);     // Yl
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                fence_async_smem();
                if (trp && !(p.flags & 64)) stamp(g, 20 + 6 * grp);
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(B_Y_READY));
                if (trp && !(p.flags & 64)) stamp(g, 21 + 6 * grp);
                store_partials();
                store_segment();
                if (itx.peek_jt() >= 0) prefetch_item(itx.peek_jt());
            }

            // per-entry math on EPC consecutive samples: z[] (TMEM columns) in, dloss/dz (TF32-rounded bits) out
            // (BATCH: 16 -- a chunk never straddles a batch segment, and the pipelined form below measured slower there)
            constexpr int EPC = BATCH ? 16 : 8;
            auto epi8 = [&](uint32_t (&z)[EPC], const float (&a)[EPC]) {
                // dmu_acc / loss_acc collect the unweighted sums; the column weight w_j (a per-thread
                // constant) is applied at the item flush.
                auto tally = [&](float gv, uint32_t zraw) -> uint32_t {
                    if (BATCH) {
                        seg_g += gv;
                        seg_gz = fmaf(gv, __uint_as_float(zraw), seg_gz);
                        return rn_bits(gv * dl);
                    }
                    dmu_acc += gv;
                    return rn_bits(gv);
                };
                if (dist == DIST_NORMAL) {
#pragma unroll
                    for (int e = 0; e < EPC; ++e) {
                        float d = fmaf(__uint_as_float(z[e]), sd, mt) - a[e];
                        d = fabsf(a[e]) < INFINITY ? d : 0.f;        // NaN / Inf => missing (ordered compare)
                        loss_acc = fmaf(d, d, loss_acc);              // (z-a)^2, halved and weighted at flush
                        z[e] = tally(d, z[e]);
                    }
                } else if (dist == DIST_BERNOULLI) {
                    // softplus(z) - a z ; sigmoid(z) - a, sharing e = exp(-|z|); a missing entry makes both NaN
                    // and is masked away once at the end
#pragma unroll
                    for (int e = 0; e < EPC; ++e) {
                        const uint32_t m = obs_mask(a[e]);
                        float z4 = fmaf(__uint_as_float(z[e]), sd, mt);
                        float ex = ex2_fast(fabsf(z4) * -1.4426950408889634f);
                        float w = 1.0f + ex;
                        float r = rcp_fast(w);
                        float sg = z4 >= 0.f ? r : ex * r;
                        float l = fmaf(-a[e], z4, fmaf(lg2_fast(w), 0.6931471805599453f, fmaxf(z4, 0.f)));
                        float gv = and_mask(sg - a[e], m);
                        loss_acc = fmaf(2.f, and_mask(l, m), loss_acc);
                        z[e] = tally(gv, z[e]);
                    }
                } else if (dist == DIST_POISSON) {
#pragma unroll
                    for (int e = 0; e < EPC; ++e) {
                        const uint32_t m = obs_mask(a[e]);
                        float z4 = fmaf(__uint_as_float(z[e]), sd, mt);
                        float ez = ex2_fast(z4 * 1.4426950408889634f);
                        float gv = and_mask(ez - a[e], m);
                        float l = and_mask(fmaf(-a[e], z4, ez), m);
                        loss_acc = fmaf(2.f, l, loss_acc);
                        z[e] = tally(gv, z[e]);
                    }
                } else {
                    // interior-threshold gradients of the ordinal models (update_noise_models): summed over the chunk
                    // here, so that nothing of it is live outside this rare branch
                    float t1s = 0.f, t2s = 0.f;
#pragma unroll
                    for (int e = 0; e < EPC; ++e) {
                        float z4 = fmaf(__uint_as_float(z[e]), sd, mt);
                        float2 lg = noise_eval_slow(dist, z4, a[e], th_range, dp.ordinal_eps, dp.hinge_margin);
                        if (THR && dp.dthr != nullptr && is_ordinal(dist)) {
                            float2 tg = threshold_grads_slow(dist, z4, a[e], th_range, dp.ordinal_eps, dp.hinge_margin);
                            t1s += tg.x; t2s += tg.y;
                        }
                        loss_acc = fmaf(2.f, lg.x, loss_acc);        // keep the common 1/2 factor at flush
                        z[e] = tally(lg.y, z[e]);
                    }
                    if (THR && dp.dthr != nullptr && is_ordinal(dist))
                        add_threshold_grads(dp.dthr, ci >> 8, t1s * wj, t2s * wj);
                }
            };


            for (int it = it0; it < it1; ++it, ++g) {
                if ((g & 1u) != (uint32_t)grp) continue;             // the other group's tile
                const bool tr = quarter == 0 && h32 == 0 && lane == 0;
                if (tr) stamp(g, 5);
                mbar_wait(bar(B_Z_FULL + rz.s), rz.ph);
                if (tr) stamp(g, 6);
                mbar_wait(bar(B_FULL_A + ra.s), ra.ph);
                tc_fence_after();
                if (tr) stamp(g, 7);
                const uint32_t zt = tm + lane_addr + TM_Z0 + 64 * rz.s + 32 * h32;
                uint8_t* abox = ag_ptr0 + ra.s * AG_BYTES + tile_box;
                if constexpr (BATCH) {
#pragma unroll 1
                for (int hh = 0; hh < 2; ++hh) {
                    uint32_t z[16];
                    float a[16];
                    TMEM_LD16(zt + 16 * hh, z);
                    // The 32-byte-atom swizzle leaves rows r and r+4 of a quarter-warp phase in the same banks.
                    // Lanes 4..7 of every 8 therefore take the two 16-byte halves of a 32-byte pair in the
                    // opposite order (address ^ 16): conflict-free 128-bit accesses, undone by register selects.
                    {
                        float4 raw[4];
#pragma unroll
                        for (int v = 0; v < 4; ++v)
                            raw[v] = *reinterpret_cast<const float4*>(abox + (chunk_off(4 * hh + v) ^ half_swap));
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            const float4 a4 = make_float4(swapped ? raw[v ^ 1].x : raw[v].x, swapped ? raw[v ^ 1].y : raw[v].y,
                                                          swapped ? raw[v ^ 1].z : raw[v].z, swapped ? raw[v ^ 1].w : raw[v].w);
                            a[4 * v] = a4.x; a[4 * v + 1] = a4.y; a[4 * v + 2] = a4.z; a[4 * v + 3] = a4.w;
                        }
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (tr && (p.flags & 64)) stamp(g, 16 + 3 * hh);
                    if (BATCH) {
                        const uint32_t id = hh == 0 ? (ids & 0xffffu) : (ids >> 16);
                        if (id != cur_b) enter_segment(id);
                    }
                    if (!(DBG && (p.ablate & 8))) epi8(z, a);
                    if (tr && (p.flags & 64)) stamp(g, 17 + 3 * hh);
                    // dloss/dz back to TMEM in place of Z (A operand of MMA3) and over the A values this thread
                    // read (MN-major A operand of MMA2): same addresses, 128-bit stores, no transposition
                    TMEM_ST16(zt + 16 * hh, z);
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const int w = v ^ 1;
                        *reinterpret_cast<uint4*>(abox + (chunk_off(4 * hh + v) ^ half_swap)) =
                            make_uint4(swapped ? z[4 * w] : z[4 * v], swapped ? z[4 * w + 1] : z[4 * v + 1],
                                       swapped ? z[4 * w + 2] : z[4 * v + 2], swapped ? z[4 * w + 3] : z[4 * v + 3]);
                    }
                    if (tr && (p.flags & 64)) stamp(g, 18 + 3 * hh);
                }
                } else {
                // The thread's 32 samples go through in four chunks of 8, software pipelined: the TMEM and shared-memory
                // loads of chunk c + 1 are issued before the math of chunk c, so their latency (TMEM read port, shared-
                // memory contention with the tensor core and TMA) is off the tile's critical path.
                // The 32-byte-atom swizzle leaves rows r and r+4 of a quarter-warp phase in the same banks.  Lanes 4..7
                // of every 8 therefore take the two 16-byte halves of a 32-byte pair in the opposite order
                // (address ^ 16): conflict-free 128-bit accesses, undone by register selects.
                uint32_t zb[2][EPC];
                float4 rawb[2][2];
                auto load_chunk = [&](int c, uint32_t (&zc)[EPC], float4 (&rc)[2]) {
                    TMEM_LD8(zt + EPC * c, zc);
#pragma unroll
                    for (int v = 0; v < 2; ++v)
                        rc[v] = *reinterpret_cast<const float4*>(abox + (chunk_off(2 * c + v) ^ half_swap));
                };
                load_chunk(0, zb[0], rawb[0]);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t (&z)[EPC] = zb[c & 1];
                    float4 (&raw)[2] = rawb[c & 1];
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (c < 3) load_chunk(c + 1, zb[(c + 1) & 1], rawb[(c + 1) & 1]);
                    float a[EPC];
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        const float4 a4 = make_float4(swapped ? raw[v ^ 1].x : raw[v].x, swapped ? raw[v ^ 1].y : raw[v].y,
                                                      swapped ? raw[v ^ 1].z : raw[v].z, swapped ? raw[v ^ 1].w : raw[v].w);
                        a[4 * v] = a4.x; a[4 * v + 1] = a4.y; a[4 * v + 2] = a4.z; a[4 * v + 3] = a4.w;
                    }
                    if (tr && (p.flags & 64) && (c & 1) == 0) stamp(g, 16 + 3 * (c >> 1));
                    if (BATCH && (c & 1) == 0) {
                        const uint32_t id = c == 0 ? (ids & 0xffffu) : (ids >> 16);
                        if (id != cur_b) enter_segment(id);
                    }
                    if (!(DBG && (p.ablate & 8))) epi8(z, a);
                    if (tr && (p.flags & 64) && (c & 1) == 1) stamp(g, 17 + 3 * (c >> 1));
                    // dloss/dz back to TMEM in place of Z (A operand of MMA3) and over the A values this thread
                    // read (MN-major A operand of MMA2): same addresses, 128-bit stores, no transposition
                    TMEM_ST8(zt + EPC * c, z);
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        const int w = v ^ 1;
                        *reinterpret_cast<uint4*>(abox + (chunk_off(2 * c + v) ^ half_swap)) =
                            make_uint4(swapped ? z[4 * w] : z[4 * v], swapped ? z[4 * w + 1] : z[4 * v + 1],
                                       swapped ? z[4 * w + 2] : z[4 * v + 2], swapped ? z[4 * w + 3] : z[4 * v + 3]);
                    }
                    if (tr && (p.flags & 64) && (c & 1) == 1) stamp(g, 18 + 3 * (c >> 1));
                }
                }
                if (tr) stamp(g, 8);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(B_G_READY + rz.s));
                store_segment();
                if (BATCH && boff >= 0 && it + 2 < it1) fetch_ids(it + 2);
                if (tr) stamp(g, 9);
                ra.next(SA); ra.next(SA);
                rz.next(SZ); rz.next(SZ);
            }
            const bool trb = quarter == 0 && h32 == 0 && lane == 0;
            if (trb && !(p.flags & 64)) stamp(g, 16 + 6 * grp);
            loss_d += (double)(loss_acc * wj);
            ItemRegs nxt;
            nxt.logsigma = nxt.mu = nxt.w = 0.f; nxt.ci = 0; nxt.boff = -1; nxt.bview = 0;
#pragma unroll
            for (int v = 0; v < 4; ++v) nxt.y[v] = make_float4(0.f, 0.f, 0.f, 0.f);

            // ---- item epilogue: dY tile out of TMEM, column sums --------------------------------------
            mbar_wait(bar(B_DY_FULL), q & 1);
            tc_fence_after();
            if (trb && !(p.flags & 64)) stamp(g, 17 + 6 * grp);
            {
                uint32_t r[16];
                TMEM_LD16(tm + lane_addr + TM_DY + 16 * c16, r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_relaxed(bar(B_DY_EMPTY));
                // next item's operands: issued after the last arrive of this item so that no fence waits on them
                if (itx.peek_jt() >= 0) load_item(itx.peek_jt(), nxt);
                // The results stay in registers and are reduced into global memory AFTER the next item's Y
                // operands have been published: REDs issued here would sit in front of that release-arrive's
                // memory fence and serialise the item boundary.
#pragma unroll
                for (int v = 0; v < 4; ++v)
                    pend_dy[v] = make_float4(gscale * __uint_as_float(r[4 * v]), gscale * __uint_as_float(r[4 * v + 1]),
                                             gscale * __uint_as_float(r[4 * v + 2]), gscale * __uint_as_float(r[4 * v + 3]));
            }
            // dmu_j = sum_i g ; dlogsigma_j = sum_i sigma_j * g  (the reference's ColScale quirk).  The four warps that
            // share a feature row (two groups x two sample halves) each store their own partial.
            if (BATCH) {
                close_segment();      // the open segment; its two REDs are deferred like the column sums
                pend_dls = dls_acc * wj * sigma;
            } else {
                pend_dls = dmu_acc * wj * sigma;
            }
            pend_dmu = dmu_acc * wj;
            pend_j = jok ? j : -1;
            if (trb && !(p.flags & 64)) stamp(g, 18 + 6 * grp);
            cur = nxt;
            ++q;
        }
        store_partials();
        store_segment();
        // data loss: sum over the epilogue warps, 0.5 factor applied here
        loss_d *= 0.5;
        for (int o = 16; o > 0; o >>= 1) loss_d += __shfl_xor_sync(0xffffffffu, loss_d, o);
        if (lane == 0) red_smem[warp] = loss_d;
    }
    tc_fence_before();
    __syncthreads();
    if (p.trace != nullptr && threadIdx.x == 0 && blockIdx.x < (unsigned)TRACE_CTAS)
        p.trace[TRACE_TILES * TRACE_EV + 2 * blockIdx.x + 1] = (long long)globaltimer_ns();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < NEPI; ++w) t += red_smem[w];
        atomicAdd(dp.scalars + SC_DATA, t);
    }
    if (warp == W_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory");
    }
}

// Operand split of X:  Xh = rna_tf32(X) (FP32 array: TF32 operand of the gradient contraction dY) and the two-term
// BF16 split Xb = [bf16(X) | bf16(X - bf16(X))] ([Mp][128] BF16: operands of the Z contraction)
// `perm` (or null): output row r holds the split of X row perm[r] (the per-order operand copies of the batch path)
__global__ void prep_operands_kernel(const float4* __restrict__ X, float4* __restrict__ Xh, uint2* __restrict__ Xb,
                                     int rows, int K4 /* Kp / 4 */, const int* stop_flag, const int32_t* __restrict__ perm) {
    if (stop_flag != nullptr && *stop_flag != 0) return;
    const size_t n4 = (size_t)rows * K4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 v;
        if (perm) {
            const int src = perm[i / K4];            // -1: padding position of the order
            v = src >= 0 ? X[(size_t)src * K4 + (i - (i / K4) * K4)] : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            v = X[i];
        }
        float4 h;
        h.x = __uint_as_float(rna_tf32(v.x)); h.y = __uint_as_float(rna_tf32(v.y));
        h.z = __uint_as_float(rna_tf32(v.z)); h.w = __uint_as_float(rna_tf32(v.w));
        const size_t row = i / K4, c4 = i - row * K4;
        Xh[row * 16 + c4] = h;                             // 16 float4 per 64-wide scratch row
        const uint32_t b01 = pack_bf16(v.x, v.y), b23 = pack_bf16(v.z, v.w);
        Xb[row * 32 + c4] = make_uint2(b01, b23);
        Xb[row * 32 + 16 + c4] = make_uint2(pack_bf16(v.x - __uint_as_float(b01 << 16), v.y - __uint_as_float(b01 & 0xffff0000u)),
                                            pack_bf16(v.z - __uint_as_float(b23 << 16), v.w - __uint_as_float(b23 & 0xffff0000u)));
    }
}

// dX[i] = sum over the sample orders of the copy's row for sample i; the copies are left zeroed for the next pass
__global__ void combine_dx_kernel(float4* __restrict__ dXo, const int32_t* __restrict__ pos, float4* __restrict__ dX,
                                  int n_orders, int M, int n_pos, int K4, const int* stop_flag) {
    if (stop_flag != nullptr && *stop_flag != 0) return;
    const size_t n4 = (size_t)M * K4;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n4; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / K4), c = (int)(idx - (size_t)i * K4);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int o = 0; o < n_orders; ++o) {
            float4* src = dXo + ((size_t)o * n_pos + pos[(size_t)o * M + i]) * K4 + c;
            const float4 v = *src;
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            *src = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        dX[idx] = acc;
    }
}

// One row of A_tc per block: the pass' feature row with its samples in the pass' order, or NaN when the
// column belongs to another pass of the same tile
__global__ void build_a_tc_kernel(const float* __restrict__ A, float* __restrict__ A_tc, int N, int lda, int n_pos,
                                  const int32_t* __restrict__ pass_feat0, const int32_t* __restrict__ pass_order,
                                  const int32_t* __restrict__ view_order, const int32_t* __restrict__ bcol_view,
                                  const int32_t* __restrict__ perm) {
    const int r = blockIdx.x, ps = r >> 7;
    const int j = pass_feat0[ps] + (r & 127);
    const int o = pass_order[ps];
    bool active = j < N;
    if (active) {
        const int bv = bcol_view[j];
        active = bv >= 0 ? view_order[bv] == o : (ps == 0 || pass_feat0[ps - 1] != pass_feat0[ps]);
    }
    float* dst = A_tc + (size_t)r * n_pos;
    const float* src = A + (size_t)(active ? j : 0) * lda;
    const int32_t* pr = perm + (size_t)o * n_pos;
    for (int p = threadIdx.x; p < n_pos; p += blockDim.x) {
        const int i = pr[p];
        dst[p] = active && i >= 0 ? src[i] : __int_as_float(0x7fc00000);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
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

// 2-D FP32 tensor [rows][cols] (cols contiguous, row pitch = pitch floats), box = box_rows x 32 floats
// 2-D BF16 tensor [rows][cols], box = box_rows x 64 elements (one 128-byte swizzle row)
bool make_map_bf16(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint32_t box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// atom32: 128B swizzle with 32-byte atoms (MN-major TF32 operands), else the standard 16-byte atoms
bool make_map(CUtensorMap* m, const float* base, uint64_t cols, uint64_t rows, uint64_t pitch, uint32_t box_rows,
              bool nan_fill, bool atom32) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {pitch * sizeof(float)};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     nan_fill ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

}  // namespace

bool tc_supported(const DataPassParams& p) {
    return p.Kp <= KK && p.Kp >= 8 && p.col_ssq == nullptr;
}

cudaError_t launch_build_a_tc(const DataPassParams& dp, const TcBatchDev& bp, float* A_tc, cudaStream_t s) {
    build_a_tc_kernel<<<bp.n_pass * 128, 256, 0, s>>>(dp.A, A_tc, dp.N, dp.lda, bp.n_pos, bp.pass_feat0, bp.pass_order,
                                                      bp.view_order, dp.bcol_view, bp.perm);
    return cudaGetLastError();
}

// Xh, Xl: [Mp][64] operand scratch owned by the handle; refreshed here when `refresh_split`
cudaError_t launch_data_pass_tc(const DataPassParams& dp_in, float* Xh, float* Xl, bool refresh_split, int precision,
                                cudaStream_t s, int n_sms, const TcBatchDev* bp, int* n_launches) {
    if (!tc_supported(dp_in) || (dp_in.n_batch_views > 0) != (bp != nullptr)) return cudaErrorInvalidValue;
    DataPassParams dp = dp_in;
    cudaError_t e = cudaSuccess;
    int launched = 0;
    const bool permuted = bp != nullptr && !bp->direct;
    int x_rows = dp.Mp;
    const int M_real = dp.M;
    if (permuted) {
        // per-order copies: operands gathered from the current X on every pass, dX gathered back afterwards.
        // The kernel sees a problem of n_pos sample POSITIONS (batches padded to multiples of 16).
        Xh = bp->Xh; Xl = bp->Xb;
        x_rows = bp->n_orders * bp->n_pos;
        refresh_split = true;
        dp.Mp = dp.lda = bp->n_pos;
        dp.M = bp->n_used;
    }
    if (refresh_split) {   // otherwise the previous epoch's update pass wrote Xh / Xl together with X
        const size_t n4 = (size_t)x_rows * dp.Kp / 4;
        prep_operands_kernel<<<(unsigned)((n4 + 255) / 256 < 1184 ? (n4 + 255) / 256 : 1184), 256, 0, s>>>(
            reinterpret_cast<const float4*>(dp.X), reinterpret_cast<float4*>(Xh), reinterpret_cast<uint2*>(Xl), x_rows,
            dp.Kp / 4, dp.stop_flag, permuted ? bp->perm : nullptr);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        ++launched;
    }
    float* dx_target = permuted ? bp->dX : dp.dX;
    const float* a_src = permuted ? bp->A_tc : dp.A;
    const int n_pass = bp ? bp->n_pass : (dp.N + BJ - 1) / BJ;
    if (bp) dp.tc_cost_cum = bp->cost_cum;

    CUtensorMap tmXb, tmXm, tmA, tmDX;
    bool ok = make_map_bf16(&tmXb, Xl, 2 * KK, x_rows, 64) && make_map(&tmXm, Xh, KK, x_rows, KK, 64, false, true) &&
              make_map(&tmA, a_src, dp.lda, permuted ? (uint64_t)n_pass * BJ : (uint64_t)dp.N, dp.lda, 128, true, true) &&
              make_map(&tmDX, dx_target, dp.Kp, x_rows, dp.Kp, 64, false, false);
    if (!ok) return cudaErrorUnknown;

    TcParams p;
    p.dp = dp;
    p.n_jt = n_pass;
    p.n_it = (dp.M + BI - 1) / BI;
    p.pass_feat0 = bp ? bp->pass_feat0 : nullptr; p.pass_order = bp ? bp->pass_order : nullptr;
    p.view_order = bp ? bp->view_order : nullptr; p.boc = bp ? bp->boc : nullptr;
    p.n_views = bp ? bp->n_views : 0; p.n_orders = bp ? bp->n_orders : 1;
    p.z_passes = precision >= 2 ? 1 : 3;
    {
        const char* ab = getenv("PMF_TC_ABLATE");      // read per launch: one process can time several settings
        p.ablate = ab ? atoi(ab) : 0;
        if (p.ablate & 4) p.z_passes = 1;
    }
    static const char* fl = getenv("PMF_TC_FLAGS");
    p.flags = fl ? atoi(fl) : 0;
    p.trace = nullptr;
    static const char* cta_times_path = getenv("PMF_TC_CTATIMES");   // per-CTA (start, end) clocks only, production kernel
    static const char* trace_path = getenv("PMF_TC_TRACE") ? getenv("PMF_TC_TRACE") : cta_times_path;
    static long long* trace_dev = nullptr;
    if (trace_path) {
        if (!trace_dev) cudaMalloc(&trace_dev, sizeof(long long) * (TRACE_TILES * TRACE_EV + 2 * TRACE_CTAS));
        cudaMemsetAsync(trace_dev, 0, sizeof(long long) * (TRACE_TILES * TRACE_EV + 2 * TRACE_CTAS), s);
        p.trace = trace_dev;
        const char* tc = getenv("PMF_TC_TRACE_CTA");
        p.trace_cta = tc ? atoi(tc) : 0;
    }
    const bool dbg = p.ablate != 0 || (p.trace != nullptr && !cta_times_path);
    const bool batch = bp != nullptr;
    const bool thr = dp.dthr != nullptr;
    auto kern = batch ? (thr ? data_pass_tc_kernel<false, true, true> : data_pass_tc_kernel<false, true, false>)
                : dbg ? data_pass_tc_kernel<true, false, false>
                      : (thr ? data_pass_tc_kernel<false, false, true> : data_pass_tc_kernel<false, false, false>);
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TOTAL);
    if (e != cudaSuccess) return e;
    const long long n_tiles = (long long)p.n_jt * p.n_it;
    int grid = n_tiles < n_sms ? (int)n_tiles : n_sms;
    kern<<<grid, NTHREADS, SMEM_TOTAL, s>>>(tmXb, tmXm, tmA, tmDX, p);
    ++launched;
    if (permuted) {
        const size_t n4 = (size_t)M_real * dp.Kp / 4;
        combine_dx_kernel<<<(unsigned)((n4 + 255) / 256 < 1184 ? (n4 + 255) / 256 : 1184), 256, 0, s>>>(
            reinterpret_cast<float4*>(bp->dX), bp->pos, reinterpret_cast<float4*>(dp_in.dX), bp->n_orders, M_real, bp->n_pos,
            dp.Kp / 4, dp.stop_flag);
        ++launched;
    }
    if (n_launches) *n_launches = launched;
    if (trace_path) {      // experiments only: dump the stamps of this launch (synchronises the stream)
        static long long host[TRACE_TILES * TRACE_EV + 2 * TRACE_CTAS];
        cudaStreamSynchronize(s);
        cudaMemcpy(host, trace_dev, sizeof host, cudaMemcpyDeviceToHost);
        if (FILE* f = fopen(trace_path, "wb")) { fwrite(host, sizeof host, 1, f); fclose(f); }
    }
    return cudaGetLastError();
}

}  // namespace pmf
