// Per-entry math of the fused data pass: noise-model loss and
// dloss/dz for one (sample, feature) entry.  Shared by the FFMA and tcgen05 kernels.
//
// The noise models live in MatFac.jl (external to the reference tree); formulas follow the
// restatement documented in DESIGN.md (per-entry math) and SURVEY.md Appendix B.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace pmf {

enum { DIST_NORMAL = 0, DIST_BERNOULLI = 1, DIST_POISSON = 2, DIST_ORDINAL3 = 3,
       DIST_BERN_SQ_HINGE = 4, DIST_ORD_SQ_HINGE3 = 5 };

__device__ __forceinline__ float sigmoidf_(float x) {
    // 1/(1+exp(-x)); exact limits at +-inf (ordinal outer thresholds)
    return 1.0f / (1.0f + __expf(-x));
}

// loss, grad of the noise model `dist` at link-space value z and datum a (finite).
// th: the range's 4 extended thresholds (only read for ordinal types).
__device__ __forceinline__ void noise_eval(int dist, float z, float a, const float* __restrict__ th,
                                           float ord_eps, float margin, float& l, float& g) {
    switch (dist) {
    case DIST_NORMAL: {
        g = z - a;
        l = 0.5f * g * g;
    } break;
    case DIST_BERNOULLI: {
        // softplus(z) - a z ; sigmoid(z) - a, sharing e = exp(-|z|)
        float e = __expf(-fabsf(z));
        float r = 1.0f / (1.0f + e);
        float s = z >= 0.f ? r : e * r;
        l = fmaxf(z, 0.f) + log1pf(e) - a * z;
        g = s - a;
    } break;
    case DIST_POISSON: {
        float ez = __expf(z);
        l = ez - a * z;
        g = ez - a;
    } break;
    case DIST_ORDINAL3: {
        int c = (int)a;                 // category 1..3
        c = c < 1 ? 1 : (c > 3 ? 3 : c);
        float lo = th[c - 1], hi = th[c];
        float sr = sigmoidf_(hi - z);
        float sl = sigmoidf_(lo - z);
        l = -__logf(sr - sl + ord_eps);
        g = 1.0f - sr - sl;
    } break;
    case DIST_BERN_SQ_HINGE: {
        float y = 2.0f * a - 1.0f;
        float h = fmaxf(0.f, 1.0f - y * z);
        l = h * h;
        g = -2.0f * y * h;
    } break;
    default: {  // DIST_ORD_SQ_HINGE3
        int c = (int)a;
        c = c < 1 ? 1 : (c > 3 ? 3 : c);
        float lo = th[c - 1], hi = th[c];
        float hl = (c > 1) ? fmaxf(0.f, lo - z + margin) : 0.f;   // -inf threshold drops
        float hr = (c < 3) ? fmaxf(0.f, z - hi + margin) : 0.f;   // +inf threshold drops
        l = hl * hl + hr * hr;
        g = 2.0f * (hr - hl);
    } break;
    }
}

// Derivatives of the two ordinal losses with respect to the INTERIOR thresholds t1 = th[1], t2 = th[2] (the outer two
// are -inf / +inf and fixed, src/fit.jl:228-242) at a finite datum a; MF.fit!'s update_noise_models (src/fit.jl:14,
// SURVEY App. D7) trains exactly these.  Category c has lo = th[c-1], hi = th[c].
__device__ __forceinline__ void noise_threshold_grads(int dist, float z, float a, const float* __restrict__ th, float ord_eps,
                                                      float margin, float& g1, float& g2) {
    int c = (int)a;
    c = c < 1 ? 1 : (c > 3 ? 3 : c);
    const float lo = th[c - 1], hi = th[c];
    float dlo, dhi;
    if (dist == DIST_ORDINAL3) {
        const float sr = sigmoidf_(hi - z), sl = sigmoidf_(lo - z);
        const float p = sr - sl + ord_eps;
        dhi = -sr * (1.0f - sr) / p;
        dlo = sl * (1.0f - sl) / p;
    } else {   // DIST_ORD_SQ_HINGE3
        const float hl = (c > 1) ? fmaxf(0.f, lo - z + margin) : 0.f;
        const float hr = (c < 3) ? fmaxf(0.f, z - hi + margin) : 0.f;
        dlo = 2.0f * hl;
        dhi = -2.0f * hr;
    }
    g1 = (c == 2 ? dlo : 0.f) + (c == 1 ? dhi : 0.f);
    g2 = (c == 3 ? dlo : 0.f) + (c == 2 ? dhi : 0.f);
}
__device__ __forceinline__ bool is_ordinal(int dist) { return dist == DIST_ORDINAL3 || dist == DIST_ORD_SQ_HINGE3; }

// Sum of one (d/dt1, d/dt2) pair per calling lane into dthr[2 * range + {0, 1}].  Called from a branch only the lanes
// with an ordinal column take: when the whole warp is there and works on one noise range (32 consecutive features: the
// normal case) the pairs are reduced with shuffles and leave as two atomics, otherwise every lane adds its own.
__device__ __forceinline__ void add_threshold_grads(float* __restrict__ dthr, int range, float g1, float g2) {
    const unsigned m = __activemask();
    bool uniform = false;
    if (m == 0xffffffffu) uniform = __all_sync(m, range == __shfl_sync(m, range, 0));
    if (uniform) {
        for (int o = 16; o > 0; o >>= 1) {
            g1 += __shfl_xor_sync(0xffffffffu, g1, o);
            g2 += __shfl_xor_sync(0xffffffffu, g2, o);
        }
        if ((threadIdx.x & 31) != 0) return;
    }
    if (g1 != 0.f) atomicAdd(dthr + 2 * range, g1);
    if (g2 != 0.f) atomicAdd(dthr + 2 * range + 1, g2);
}

__device__ __forceinline__ bool is_observed(float a) {
    // missing data is NaN-encoded; non-finite values never contribute (bit test, immune to
    // fast-math comparisons)
    return (__float_as_uint(a) & 0x7f800000u) != 0x7f800000u;
}

__device__ __forceinline__ double block_reduce_sum_double(double v, double* smem_warp) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) smem_warp[warp] = v;
    __syncthreads();
    double t = 0.0;
    if (warp == 0) {
        int nw = (blockDim.x + 31) >> 5;
        t = lane < nw ? smem_warp[lane] : 0.0;
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    }
    return t;   // valid in thread 0
}

}  // namespace pmf
