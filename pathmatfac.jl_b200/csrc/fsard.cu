// Feature-set ARD outer step on the device: update_A! / update_A_inner!
// (src/featureset_ard.jl:214-294) = projected AdaGrad-ISTA (src/optimizers.jl:46-62) on
// gamma_normal_loss (src/featureset_ard.jl:154-186), with the reference's best-so-far /
// termination-counter bookkeeping evaluated on the device (host polls a flag).
//
//   beta      = beta0 (v0 + A'S)                         K x N_v   (stored [N_v][Kp])
//   loss      = -sum_j a_j sum_k log beta + sum_j (a_j+.5) sum_k log(beta + .5 Y^2) - calibration
//   grad_AtS  = beta0 (-a_j / beta + (a_j+.5) / (beta + .5 Y^2))
//   grad_A    = S grad_AtS'                              L x K
#include "../../include/pmf.h"
#include "pmf_host.h"
#include "pmf_epilogue.cuh"

#include <algorithm>
#include <cstring>
#include <vector>

namespace {

struct FsCtrl {
    double best_loss;
    double cur_smooth;     // gamma-normal part of the loss at the current A
    double cur_reg;        // lambda-weighted L1 part at the current A
    double calib;          // calibration constant (independent of A)
    int term_count;
    int stop;
    int epochs;
    int improved;
    int hit_term;
};

constexpr int FT = 256;

// one thread per (feature j, factor k): beta, loss term and dLoss/d(A'S)
__global__ void __launch_bounds__(FT) fs_beta_kernel(int Nv, int K, int Kp, const int32_t* __restrict__ st_rp,
                                                     const int32_t* __restrict__ st_col, const float* __restrict__ st_val,
                                                     const float* __restrict__ A, const float* __restrict__ Yv,
                                                     const float* __restrict__ alpha, float beta0, float v0,
                                                     float* __restrict__ beta_out, float* __restrict__ G, FsCtrl* c,
                                                     int calib_pass) {
    if (c->stop) return;
    __shared__ double red[FT / 32];
    double lsum = 0.0;
    const size_t total = (size_t)Nv * Kp;
    for (size_t idx = (size_t)blockIdx.x * FT + threadIdx.x; idx < total; idx += (size_t)gridDim.x * FT) {
        const int j = (int)(idx / Kp), k = (int)(idx - (size_t)j * Kp);
        if (k >= K) continue;
        float ats = 0.f;
        for (int e = st_rp[j]; e < st_rp[j + 1]; ++e) ats = fmaf(st_val[e], A[(size_t)st_col[e] * Kp + k], ats);
        const float beta = beta0 * (v0 + ats);
        const float y = Yv[(size_t)j * Kp + k];
        const float a = alpha[j], ap5 = a + 0.5f;
        const float den = beta + 0.5f * y * y;
        if (calib_pass) {
            // sum_jk log(|Y| + 1e-9) + sum_j [(a+.5) log(a+.5) - a log a]   (the latter once per j)
            lsum += (double)logf(fabsf(y) + 1e-9f);
            if (k == 0) lsum += (double)(ap5 * logf(ap5) - a * logf(a));
        } else {
            lsum += (double)(-a * logf(beta) + ap5 * logf(den));
            G[idx] = beta0 * (-a / beta + ap5 / den);
            beta_out[idx] = beta;
        }
    }
    double tot = pmf::block_reduce_sum_double(lsum, red);
    if (threadIdx.x == 0) atomicAdd(calib_pass ? &c->calib : &c->cur_smooth, tot);
}

// one thread per (set l, factor k): grad_A, ISTA step, L1 term of the new A
__global__ void __launch_bounds__(FT) fs_ista_kernel(int L, int K, int Kp, const int32_t* __restrict__ s_rp,
                                                     const int32_t* __restrict__ s_col, const float* __restrict__ s_val,
                                                     const float* __restrict__ G, float* __restrict__ A,
                                                     float* __restrict__ ssq, const float* __restrict__ lambda, float lr,
                                                     FsCtrl* c) {
    if (c->stop) return;
    __shared__ double red[FT / 32];
    double lsum = 0.0;
    const size_t total = (size_t)L * Kp;
    for (size_t idx = (size_t)blockIdx.x * FT + threadIdx.x; idx < total; idx += (size_t)gridDim.x * FT) {
        const int l = (int)(idx / Kp), k = (int)(idx - (size_t)l * Kp);
        if (k >= K) continue;
        float g = 0.f;
        for (int e = s_rp[l]; e < s_rp[l + 1]; ++e) g = fmaf(s_val[e], G[(size_t)s_col[e] * Kp + k], g);
        float q = ssq[idx] + g * g;
        ssq[idx] = q;
        const float eta = lr / sqrtf(q);
        float a = A[idx] - eta * g;
        a = fmaxf(a, 0.f);
        a = fmaxf(fabsf(a) - lambda[k] * eta, 0.f);
        A[idx] = a;
        lsum += (double)(lambda[k] * fabsf(a));
    }
    double tot = pmf::block_reduce_sum_double(lsum, red);
    if (threadIdx.x == 0) atomicAdd(&c->cur_reg, tot);
}

__global__ void fs_begin_epoch(FsCtrl* c) {
    if (c->stop) return;
    c->cur_smooth = 0.0;
    c->cur_reg = 0.0;
}

// best-so-far / termination counter of update_A_inner! (featureset_ard.jl:237-263)
__global__ void fs_control(FsCtrl* c, int epoch, int term_iter, double atol, int initial) {
    if (c->stop) return;
    const double loss = c->cur_smooth - c->calib + c->cur_reg;
    if (initial) {
        c->best_loss = loss;
        c->improved = 1;
        return;
    }
    c->epochs = epoch;
    if (loss < c->best_loss) {
        double d = c->best_loss - loss;
        c->best_loss = loss;
        c->improved = 1;
        c->term_count = d > atol ? 0 : c->term_count + 1;
    } else {
        c->improved = 0;
        c->term_count += 1;
    }
    if (c->term_count >= term_iter) {
        c->hit_term = 1;
        c->stop = 2;     // stop after A_best has been refreshed by fs_keep_best
    }
}

__global__ void fs_keep_best(const float* __restrict__ A, float* __restrict__ Abest, size_t n, FsCtrl* c) {
    if (c->stop == 1) return;
    if (c->improved)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) Abest[i] = A[i];
}

__global__ void fs_finish_epoch(FsCtrl* c) {
    if (c->stop == 2) c->stop = 1;
}

template <class T>
cudaError_t dalloc(T** p, size_t n) {
    *p = nullptr;
    return cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T));
}

}  // namespace

extern "C" int pmf_fsard_update_A(pmf_handle h, int32_t col_start, int32_t col_stop, int32_t L, const int32_t* S_rp,
                                  const int32_t* S_col, const float* S_val, float* A_host, float* ssq_host,
                                  const float* lambda_K, float lr, float alpha0, float v0, int32_t max_epochs,
                                  int32_t term_iter, float atol, double* best_loss, int32_t* epochs_run) {
    if (!h) return PMF_ERR_ARG;
    auto bad = [&](int code, const char* m) { h->err = m; return code; };
    if (h->cuda_failed) return PMF_ERR_CUDA;
    if (cudaSetDevice(h->dims.device) != cudaSuccess) return bad(PMF_ERR_CUDA, "cudaSetDevice failed");
    SideReg& r = h->reg[1];
    if (!r.ard_alpha || !r.ard_beta_full) return bad(PMF_ERR_STATE, "pmf_set_reg_fsard(which=1) must be called first");
    if (col_start < 0 || col_stop > h->N || col_start >= col_stop || L <= 0 || !S_rp || !A_host || !ssq_host || !lambda_K)
        return bad(PMF_ERR_ARG, "bad FSARD update arguments");
    const int Nv = col_stop - col_start, K = h->K, Kp = h->Kp;
    const int nnz = S_rp[L];
    // S^T (N_v rows) on the host
    std::vector<int32_t> t_rp(Nv + 1, 0), t_col(std::max(nnz, 1));
    std::vector<float> t_val(std::max(nnz, 1));
    for (int e = 0; e < nnz; ++e) {
        if (S_col[e] < 0 || S_col[e] >= Nv) return bad(PMF_ERR_ARG, "feature-set column index out of range");
        t_rp[S_col[e] + 1]++;
    }
    for (int j = 0; j < Nv; ++j) t_rp[j + 1] += t_rp[j];
    {
        std::vector<int32_t> fill(t_rp.begin(), t_rp.end() - 1);
        for (int l = 0; l < L; ++l)
            for (int e = S_rp[l]; e < S_rp[l + 1]; ++e) {
                int pos = fill[S_col[e]]++;
                t_col[pos] = l;
                t_val[pos] = S_val[e];
            }
    }
    cudaStream_t s = h->stream;
    int32_t *d_srp, *d_scol, *d_trp, *d_tcol;
    float *d_sval, *d_tval, *d_A, *d_Abest, *d_ssq, *d_G, *d_lam;
    FsCtrl* d_c;
    bool ok = dalloc(&d_srp, L + 1) == cudaSuccess && dalloc(&d_scol, nnz) == cudaSuccess && dalloc(&d_sval, nnz) == cudaSuccess &&
              dalloc(&d_trp, Nv + 1) == cudaSuccess && dalloc(&d_tcol, nnz) == cudaSuccess && dalloc(&d_tval, nnz) == cudaSuccess &&
              dalloc(&d_A, (size_t)L * Kp) == cudaSuccess && dalloc(&d_Abest, (size_t)L * Kp) == cudaSuccess &&
              dalloc(&d_ssq, (size_t)L * Kp) == cudaSuccess && dalloc(&d_G, (size_t)Nv * Kp) == cudaSuccess &&
              dalloc(&d_lam, Kp) == cudaSuccess && dalloc(&d_c, 1) == cudaSuccess;
    auto cleanup = [&]() {
        cudaFree(d_srp); cudaFree(d_scol); cudaFree(d_sval); cudaFree(d_trp); cudaFree(d_tcol); cudaFree(d_tval);
        cudaFree(d_A); cudaFree(d_Abest); cudaFree(d_ssq); cudaFree(d_G); cudaFree(d_lam); cudaFree(d_c);
    };
    if (!ok) { cleanup(); return bad(PMF_ERR_ALLOC, "device allocation failed"); }
    cudaMemcpyAsync(d_srp, S_rp, (size_t)(L + 1) * 4, cudaMemcpyHostToDevice, s);
    cudaMemcpyAsync(d_trp, t_rp.data(), (size_t)(Nv + 1) * 4, cudaMemcpyHostToDevice, s);
    if (nnz > 0) {
        cudaMemcpyAsync(d_scol, S_col, (size_t)nnz * 4, cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(d_sval, S_val, (size_t)nnz * 4, cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(d_tcol, t_col.data(), (size_t)nnz * 4, cudaMemcpyHostToDevice, s);
        cudaMemcpyAsync(d_tval, t_val.data(), (size_t)nnz * 4, cudaMemcpyHostToDevice, s);
    }
    // A is zeroed first like update_A! (featureset_ard.jl:286); ssq_grad is the optimiser's state
    cudaMemsetAsync(d_A, 0, (size_t)L * Kp * 4, s);
    cudaMemsetAsync(d_Abest, 0, (size_t)L * Kp * 4, s);
    cudaMemsetAsync(d_ssq, 0, (size_t)L * Kp * 4, s);
    cudaMemcpy2DAsync(d_ssq, (size_t)Kp * 4, ssq_host, (size_t)K * 4, (size_t)K * 4, L, cudaMemcpyHostToDevice, s);
    std::vector<float> lam(Kp, 0.f);
    std::memcpy(lam.data(), lambda_K, (size_t)K * 4);
    cudaMemcpyAsync(d_lam, lam.data(), (size_t)Kp * 4, cudaMemcpyHostToDevice, s);
    FsCtrl c0;
    std::memset(&c0, 0, sizeof c0);
    cudaMemcpyAsync(d_c, &c0, sizeof c0, cudaMemcpyHostToDevice, s);

    const float beta0 = alpha0 - 1.0f;
    const float* Yv = h->Y + (size_t)col_start * Kp;
    const float* alpha = r.ard_alpha + col_start;
    float* beta_out = r.ard_beta_full + (size_t)col_start * Kp;
    const int gb = (int)std::min<size_t>(((size_t)Nv * Kp + FT - 1) / FT, 148 * 8);
    const int ga = (int)std::min<size_t>(((size_t)L * Kp + FT - 1) / FT, 148 * 8);

    // calibration constant, then loss / gradient at A = 0
    fs_beta_kernel<<<gb, FT, 0, s>>>(Nv, K, Kp, d_trp, d_tcol, d_tval, d_A, Yv, alpha, beta0, v0, beta_out, d_G, d_c, 1);
    fs_beta_kernel<<<gb, FT, 0, s>>>(Nv, K, Kp, d_trp, d_tcol, d_tval, d_A, Yv, alpha, beta0, v0, beta_out, d_G, d_c, 0);
    fs_control<<<1, 1, 0, s>>>(d_c, 0, term_iter, (double)atol, 1);
    h->launches += 3;
    FsCtrl hc;
    for (int epoch = 1; epoch <= max_epochs; ++epoch) {
        fs_begin_epoch<<<1, 1, 0, s>>>(d_c);
        fs_ista_kernel<<<ga, FT, 0, s>>>(L, K, Kp, d_srp, d_scol, d_sval, d_G, d_A, d_ssq, d_lam, lr, d_c);
        fs_beta_kernel<<<gb, FT, 0, s>>>(Nv, K, Kp, d_trp, d_tcol, d_tval, d_A, Yv, alpha, beta0, v0, beta_out, d_G, d_c, 0);
        fs_control<<<1, 1, 0, s>>>(d_c, epoch, term_iter, (double)atol, 0);
        fs_keep_best<<<ga, FT, 0, s>>>(d_A, d_Abest, (size_t)L * Kp, d_c);
        fs_finish_epoch<<<1, 1, 0, s>>>(d_c);
        h->launches += 6;
        if (epoch % 16 == 0 || epoch == max_epochs) {
            cudaMemcpyAsync(&hc, d_c, sizeof hc, cudaMemcpyDeviceToHost, s);
            if (cudaStreamSynchronize(s) != cudaSuccess) { cleanup(); h->cuda_failed = true; return bad(PMF_ERR_CUDA, "FSARD update failed"); }
            if (hc.stop) break;
        }
    }
    cudaMemcpyAsync(&hc, d_c, sizeof hc, cudaMemcpyDeviceToHost, s);
    cudaStreamSynchronize(s);
    // A .= A_best ; beta[:, cr] = beta0 (v0 + A'S)   (featureset_ard.jl:272, :292)
    c0 = hc;
    c0.stop = 0; c0.cur_smooth = 0.0;
    cudaMemcpyAsync(d_c, &c0, sizeof c0, cudaMemcpyHostToDevice, s);
    fs_beta_kernel<<<gb, FT, 0, s>>>(Nv, K, Kp, d_trp, d_tcol, d_tval, d_Abest, Yv, alpha, beta0, v0, beta_out, d_G, d_c, 0);
    h->launches += 1;
    cudaMemcpy2DAsync(A_host, (size_t)K * 4, d_Abest, (size_t)Kp * 4, (size_t)K * 4, L, cudaMemcpyDeviceToHost, s);
    cudaMemcpy2DAsync(ssq_host, (size_t)K * 4, d_ssq, (size_t)Kp * 4, (size_t)K * 4, L, cudaMemcpyDeviceToHost, s);
    cudaError_t e = cudaStreamSynchronize(s);
    cleanup();
    if (e != cudaSuccess) { h->cuda_failed = true; return bad(PMF_ERR_CUDA, cudaGetErrorString(e)); }
    if (best_loss) *best_loss = hc.best_loss;
    if (epochs_run) *epochs_run = hc.epochs;
    return PMF_OK;
}
