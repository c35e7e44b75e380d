// Fused data pass, FP32 CUDA-core (FFMA) tile kernel.
//
// One CTA owns a 64-feature tile and a chunk of samples.  Per 64x64 tile of A it
//   (1) forms Z = X_t' Y_t in registers (4x4 micro-tiles, K-deep FFMA from shared memory),
//   (2) applies ColScale / BatchScale / ColShift / BatchShift (reference order,
//       src/layers.jl:221-253), evaluates the per-assay noise loss and dL/dz with the NaN mask,
//       and accumulates the column / batch parameter gradients (src/layers.jl:34-90,
//       src/batch_array.jl:132-212 incl. the ColScale quirk logsigma_bar = sum sigma*Gbar),
//   (3) contracts G0 with X_t (-> dY, kept in registers across the sample loop) and with Y_t
//       (-> dX, vector atomics to global).
// Z and dL/dZ never touch HBM; A is read exactly once.  Products are exact FP32, so this
// kernel is also the on-device reference for the tcgen05 path.
#include "pmf_internal.h"
#include "pmf_epilogue.cuh"

namespace pmf {

namespace {
constexpr int TJ = 64;    // features per tile
constexpr int TI = 64;    // samples per tile
constexpr int NT = 256;   // threads
constexpr int GP = 68;    // pitch of the G0 tile in shared memory (bank-conflict free float4 rows)

__device__ __forceinline__ void red_add4(float* addr, float4 v) {
    atomicAdd(reinterpret_cast<float4*>(addr), v);   // RED.E.ADD.F32x4 on sm_90+
}

template <int KG, bool STATS>
__global__ void __launch_bounds__(NT, (KG == 1) ? 2 : 1) data_pass_ffma_kernel(DataPassParams p) {
    if (p.stop_flag != nullptr && *p.stop_flag != 0) return;
    extern __shared__ __align__(16) float smem[];
    __shared__ double red_smem[NT / 32];

    const int Kp = p.Kp, KP4 = Kp + 4, K4 = Kp >> 2;
    const bool has_batch = p.n_batch_views > 0;
    const int nbm = has_batch ? p.nb_max : 0;
    float* Ys = smem;                 // [TJ][KP4]
    float* Xs = Ys + TJ * KP4;        // [TI][KP4]
    float* Gs = Xs + TI * KP4;        // [TJ][GP]
    float* Bth = Gs + TJ * GP;        // [TJ][nbm] theta-gradient partials
    float* Bld = Bth + TJ * nbm;      // [TJ][nbm] logdelta-gradient partials (sum z1*g)

    const int tid = threadIdx.x, tj = tid >> 4, ti = tid & 15;
    const int j0 = blockIdx.x * TJ;
    const int n_itiles = (p.M + TI - 1) / TI;
    const int it_begin = (int)((long long)n_itiles * blockIdx.y / gridDim.y);
    const int it_end = (int)((long long)n_itiles * (blockIdx.y + 1) / gridDim.y);
    if (it_begin >= it_end) return;

    // ---- feature tile of Y (resident for the whole sample loop) -------------------------
    for (int idx = tid; idx < TJ * K4; idx += NT) {
        int row = idx / K4, c4 = idx - row * K4;
        float4 v = reinterpret_cast<const float4*>(p.Y + (size_t)(j0 + row) * Kp)[c4];
        *reinterpret_cast<float4*>(&Ys[row * KP4 + 4 * c4]) = v;
    }
    if (has_batch)
        for (int idx = tid; idx < 2 * TJ * nbm; idx += NT) Bth[idx] = 0.f;

    // ---- per-column constants of this thread's 4 features --------------------------------
    float sig[4], muv[4], wv[4];
    int dist[4], boff[4], bview[4];
    const float* thp[4];
    bool jok[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        int j = j0 + tj + 16 * a;
        jok[a] = j < p.N;
        int jj = jok[a] ? j : 0;
        sig[a] = __expf(p.logsigma[jj]);
        muv[a] = p.mu[jj];
        wv[a] = jok[a] ? p.weight[jj] : 0.f;
        int ci = p.colinfo[jj];
        dist[a] = ci & 0xff;
        thp[a] = p.thresholds + 4 * (ci >> 8);
        boff[a] = (has_batch && jok[a]) ? p.bcol_off[jj] : -1;
        bview[a] = boff[a] >= 0 ? p.bcol_view[jj] : 0;
    }

    float dYacc[KG][4][4];
#pragma unroll
    for (int g = 0; g < KG; ++g)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) dYacc[g][a][c] = 0.f;
    float dmu_acc[4] = {0.f, 0.f, 0.f, 0.f}, dls_acc[4] = {0.f, 0.f, 0.f, 0.f};
    float ssq_acc[4] = {0.f, 0.f, 0.f, 0.f}, cnt_acc[4] = {0.f, 0.f, 0.f, 0.f}, sq_acc[4] = {0.f, 0.f, 0.f, 0.f};
    float loss_acc = 0.f;
    double loss_d = 0.0;
    float thr_acc[4][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};   // interior ordinal thresholds (p.dthr)

    for (int it = it_begin; it < it_end; ++it) {
        const int i0 = it * TI;
        __syncthreads();   // previous tile's readers of Xs / Gs are done
        for (int idx = tid; idx < TI * K4; idx += NT) {
            int row = idx / K4, c4 = idx - row * K4;
            float4 v = reinterpret_cast<const float4*>(p.X + (size_t)(i0 + row) * Kp)[c4];
            *reinterpret_cast<float4*>(&Xs[row * KP4 + 4 * c4]) = v;
        }
        // A entries of this thread's micro-tile, issued before the K loop (latency hidden)
        float av[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float* Arow = p.A + (size_t)(j0 + tj + 16 * a) * p.lda + i0 + ti;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                int i = i0 + ti + 16 * b;
                av[a][b] = (jok[a] && i < p.M) ? __ldg(Arow + 16 * b) : __int_as_float(0x7fc00000);
            }
        }
        __syncthreads();

        // ---- (1) Z micro-tile ---------------------------------------------------------------
        float z[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) z[a][b] = 0.f;
        for (int k4 = 0; k4 < K4; ++k4) {
            float4 y4[4], x4[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) y4[a] = *reinterpret_cast<const float4*>(&Ys[(tj + 16 * a) * KP4 + 4 * k4]);
#pragma unroll
            for (int b = 0; b < 4; ++b) x4[b] = *reinterpret_cast<const float4*>(&Xs[(ti + 16 * b) * KP4 + 4 * k4]);
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    z[a][b] = fmaf(y4[a].x, x4[b].x, z[a][b]);
                    z[a][b] = fmaf(y4[a].y, x4[b].y, z[a][b]);
                    z[a][b] = fmaf(y4[a].z, x4[b].z, z[a][b]);
                    z[a][b] = fmaf(y4[a].w, x4[b].w, z[a][b]);
                }
        }

        // ---- (2) layers, loss, dL/dz, parameter-gradient partials ------------------------------
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int i = i0 + ti + 16 * b;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                float aval = av[a][b];
                float g0 = 0.f;
                if (is_observed(aval)) {
                    float z1 = z[a][b] * sig[a];
                    float delta = 1.f, th = 0.f;
                    int bidx = 0;
                    if (boff[a] >= 0) {
                        bidx = p.batch_of_sample[(size_t)bview[a] * p.Mp + i];
                        delta = __expf(p.logdelta[boff[a] + bidx]);
                        th = p.theta[boff[a] + bidx];
                    }
                    float z2 = z1 * delta;
                    float z4 = z2 + muv[a] + th;
                    float l, g;
                    noise_eval(dist[a], z4, aval, thp[a], p.ordinal_eps, p.hinge_margin, l, g);
                    if (!STATS && p.dthr != nullptr && is_ordinal(dist[a])) {
                        float t1, t2;
                        noise_threshold_grads(dist[a], z4, aval, thp[a], p.ordinal_eps, p.hinge_margin, t1, t2);
                        thr_acc[a][0] = fmaf(wv[a], t1, thr_acc[a][0]);
                        thr_acc[a][1] = fmaf(wv[a], t2, thr_acc[a][1]);
                    }
                    l *= wv[a];
                    g *= wv[a];
                    loss_acc += l;
                    if (STATS) {
                        // statistics pass of the staging code: sum (dl/dz)^2, count of finite entries, squared
                        // error in link space, and the last two segmented by batch (ba_map, src/batch_array.jl:320)
                        const float d = aval - z4;
                        ssq_acc[a] += g * g;
                        cnt_acc[a] += 1.f;
                        sq_acc[a] += d * d;
                        if (boff[a] >= 0) {
                            int lj = tj + 16 * a;
                            atomicAdd(&Bth[lj * nbm + bidx], 1.f);
                            atomicAdd(&Bld[lj * nbm + bidx], d * d);
                        }
                    } else {
                        dmu_acc[a] += g;
                        float g1 = g * delta;
                        g0 = g1 * sig[a];
                        dls_acc[a] += g0;          // ColScale quirk: sigma_j * Gbar_ij, no Z factor
                        if (boff[a] >= 0) {
                            int lj = tj + 16 * a;
                            atomicAdd(&Bth[lj * nbm + bidx], g);
                            atomicAdd(&Bld[lj * nbm + bidx], g * z2);   // exp(logdelta) * z1 * g
                        }
                    }
                }
                if (!STATS) Gs[(tj + 16 * a) * GP + ti + 16 * b] = g0;
            }
        }
        loss_d += (double)loss_acc;
        loss_acc = 0.f;
        if (STATS) continue;
        __syncthreads();

        // ---- (3a) dY[j][k] += sum_i G0[j][i] X[i][k] -------------------------------------------
#pragma unroll 1
        for (int i4 = 0; i4 < TI / 4; ++i4) {
            float4 g4[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) g4[a] = *reinterpret_cast<const float4*>(&Gs[(tj + 16 * a) * GP + 4 * i4]);
#pragma unroll
            for (int g = 0; g < KG; ++g) {
                int kb = 4 * ti + 64 * g;
                if (kb < Kp) {
#pragma unroll
                    for (int ii = 0; ii < 4; ++ii) {
                        float4 x4 = *reinterpret_cast<const float4*>(&Xs[(4 * i4 + ii) * KP4 + kb]);
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            float gv = ii == 0 ? g4[a].x : (ii == 1 ? g4[a].y : (ii == 2 ? g4[a].z : g4[a].w));
                            dYacc[g][a][0] = fmaf(gv, x4.x, dYacc[g][a][0]);
                            dYacc[g][a][1] = fmaf(gv, x4.y, dYacc[g][a][1]);
                            dYacc[g][a][2] = fmaf(gv, x4.z, dYacc[g][a][2]);
                            dYacc[g][a][3] = fmaf(gv, x4.w, dYacc[g][a][3]);
                        }
                    }
                }
            }
        }

        // ---- (3b) dX[i][k] += sum_j G0[j][i] Y[j][k]   (this tile's contribution) -------------
        {
            const int tk = tid & 15, tq = tid >> 4;   // i = 4*tq + b, k = 4*tk + 64*g + c
#pragma unroll
            for (int g = 0; g < KG; ++g) {
                int kb = 4 * tk + 64 * g;
                if (kb >= Kp) continue;
                float dx[4][4];
#pragma unroll
                for (int b = 0; b < 4; ++b)
#pragma unroll
                    for (int c = 0; c < 4; ++c) dx[b][c] = 0.f;
#pragma unroll 8
                for (int j = 0; j < TJ; ++j) {
                    float4 g4 = *reinterpret_cast<const float4*>(&Gs[j * GP + 4 * tq]);
                    float4 y4 = *reinterpret_cast<const float4*>(&Ys[j * KP4 + kb]);
                    dx[0][0] = fmaf(g4.x, y4.x, dx[0][0]); dx[0][1] = fmaf(g4.x, y4.y, dx[0][1]);
                    dx[0][2] = fmaf(g4.x, y4.z, dx[0][2]); dx[0][3] = fmaf(g4.x, y4.w, dx[0][3]);
                    dx[1][0] = fmaf(g4.y, y4.x, dx[1][0]); dx[1][1] = fmaf(g4.y, y4.y, dx[1][1]);
                    dx[1][2] = fmaf(g4.y, y4.z, dx[1][2]); dx[1][3] = fmaf(g4.y, y4.w, dx[1][3]);
                    dx[2][0] = fmaf(g4.z, y4.x, dx[2][0]); dx[2][1] = fmaf(g4.z, y4.y, dx[2][1]);
                    dx[2][2] = fmaf(g4.z, y4.z, dx[2][2]); dx[2][3] = fmaf(g4.z, y4.w, dx[2][3]);
                    dx[3][0] = fmaf(g4.w, y4.x, dx[3][0]); dx[3][1] = fmaf(g4.w, y4.y, dx[3][1]);
                    dx[3][2] = fmaf(g4.w, y4.z, dx[3][2]); dx[3][3] = fmaf(g4.w, y4.w, dx[3][3]);
                }
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    int i = i0 + 4 * tq + b;
                    if (i < p.M)
                        red_add4(p.dX + (size_t)i * Kp + kb, make_float4(dx[b][0], dx[b][1], dx[b][2], dx[b][3]));
                }
            }
        }
    }

    // ---- flush per-CTA accumulators ------------------------------------------------------------
    if (!STATS) {
#pragma unroll
        for (int g = 0; g < KG; ++g) {
            int kb = 4 * ti + 64 * g;
            if (kb >= Kp) continue;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                if (!jok[a]) continue;
                float* dst = p.dY + (size_t)(j0 + tj + 16 * a) * Kp + kb;
                float4 v = make_float4(dYacc[g][a][0], dYacc[g][a][1], dYacc[g][a][2], dYacc[g][a][3]);
                if (gridDim.y == 1) *reinterpret_cast<float4*>(dst) = v;
                else red_add4(dst, v);
            }
        }
    }
    // column sums: reduce over the 16 lanes (ti) that share a feature row
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        float v0 = STATS ? ssq_acc[a] : dmu_acc[a];
        float v1 = STATS ? cnt_acc[a] : dls_acc[a];
        float v2 = STATS ? sq_acc[a] : 0.f;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            v0 += __shfl_xor_sync(0xffffffffu, v0, o);
            v1 += __shfl_xor_sync(0xffffffffu, v1, o);
            v2 += __shfl_xor_sync(0xffffffffu, v2, o);
        }
        if (ti == 0 && jok[a]) {
            int j = j0 + tj + 16 * a;
            if (STATS) {
                atomicAdd(p.col_ssq + j, v0);
                atomicAdd(p.col_cnt + j, v1);
                if (p.col_sqerr) atomicAdd(p.col_sqerr + j, v2);
            } else {
                atomicAdd(p.dmu + j, v0);
                atomicAdd(p.dlogsigma + j, v1);
            }
        }
    }
    if (!STATS && p.dthr != nullptr) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int range = (int)((thp[a] - p.thresholds) >> 2);
            if (thr_acc[a][0] != 0.f) atomicAdd(p.dthr + 2 * range, thr_acc[a][0]);
            if (thr_acc[a][1] != 0.f) atomicAdd(p.dthr + 2 * range + 1, thr_acc[a][1]);
        }
    }
    if (has_batch) {      // gradients of theta / logdelta, or (statistics pass) per-batch counts / squared errors
        __syncthreads();
        for (int idx = tid; idx < TJ * nbm; idx += NT) {
            int lj = idx / nbm, b = idx - lj * nbm;
            int j = j0 + lj;
            if (j >= p.N) continue;
            int off = p.bcol_off[j];
            if (off < 0 || b >= p.bcol_nb[j]) continue;
            float gth = Bth[idx], gld = Bld[idx];
            if (gth != 0.f) atomicAdd(p.dtheta + off + b, gth);
            if (gld != 0.f) atomicAdd(p.dlogdelta + off + b, gld);
        }
    }
    double tot = block_reduce_sum_double(loss_d, red_smem);
    if (tid == 0 && !STATS) atomicAdd(p.scalars + SC_DATA, tot);
}

template <int KG, bool STATS>
cudaError_t launch_impl(const DataPassParams& p, cudaStream_t s, int n_sms) {
    const int nbm = p.n_batch_views > 0 ? p.nb_max : 0;
    size_t smem = sizeof(float) * ((size_t)(TJ + TI) * (p.Kp + 4) + (size_t)TJ * GP + 2 * (size_t)TJ * nbm);
    auto kern = data_pass_ffma_kernel<KG, STATS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int n_jt = (p.N + TJ - 1) / TJ;
    int n_it = (p.M + TI - 1) / TI;
    int chunks = p.sample_chunks;
    if (chunks <= 0) {
        // aim for >= ~6 CTAs per SM in total so the tail wave is short; at least 4 tiles a chunk
        chunks = (6 * n_sms + n_jt - 1) / n_jt;
        int max_chunks = n_it / 4 > 0 ? n_it / 4 : 1;
        if (chunks > max_chunks) chunks = max_chunks;
        if (chunks < 1) chunks = 1;
    }
    dim3 grid(n_jt, chunks);
    kern<<<grid, NT, smem, s>>>(p);
    return cudaGetLastError();
}
}  // namespace

cudaError_t launch_data_pass_ffma(const DataPassParams& p, cudaStream_t s, int n_sms) {
    const bool stats = p.col_ssq != nullptr;
    const int kg = (p.Kp + 63) / 64;
    if (kg == 1) return stats ? launch_impl<1, true>(p, s, n_sms) : launch_impl<1, false>(p, s, n_sms);
    if (kg == 2) return stats ? launch_impl<2, true>(p, s, n_sms) : launch_impl<2, false>(p, s, n_sms);
    if (kg <= 4) return stats ? launch_impl<4, true>(p, s, n_sms) : launch_impl<4, false>(p, s, n_sms);
    return cudaErrorInvalidValue;
}

}  // namespace pmf
