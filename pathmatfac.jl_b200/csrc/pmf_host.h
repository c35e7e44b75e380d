// Host-side handle behind the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/pmf.h"
#include "pmf_internal.h"

struct BatchView {
    int col_start, col_stop, n_batches;
    int64_t offset;   // into the flattened [column][batch] tables
};

struct DevCsr {
    int32_t* rowptr = nullptr;
    int32_t* col = nullptr;
    float* val = nullptr;
    int64_t* rowptr_base = nullptr;
    int64_t* nnz_base = nullptr;
    void free_all() {
        cudaFree(rowptr); cudaFree(col); cudaFree(val); cudaFree(rowptr_base); cudaFree(nnz_base);
        rowptr = col = nullptr; val = nullptr; rowptr_base = nnz_base = nullptr;
    }
};

struct DevNetwork {
    bool present = false;
    DevCsr AA, AB, BB, ABt;
    int32_t* nv = nullptr;
    int64_t* virt_base = nullptr;
    float* u = nullptr;
    float* work = nullptr;
    float* yt = nullptr;      // [K][ld] transposed factor matrix / per-factor gradient (scratch of the pass)
    float* gt = nullptr;
    int ld = 0, nv_max = 0;
    int64_t nv_total = 0;
    float p = 1.f, rtol = 0.f, atol = 0.f;
    int itmax = 0;
    void free_all() {
        AA.free_all(); AB.free_all(); BB.free_all(); ABt.free_all();
        cudaFree(nv); cudaFree(virt_base); cudaFree(u); cudaFree(work); cudaFree(yt); cudaFree(gt);
        nv = nullptr; virt_base = nullptr; u = work = yt = gt = nullptr;
        present = false; nv_total = 0;
    }
};

struct SideReg {
    float* l2_w = nullptr;
    int32_t* group_id = nullptr;
    float* group_w = nullptr;
    uint8_t* l1_mask = nullptr;
    float* l1_w = nullptr;
    float* ard_alpha = nullptr;
    float* ard_beta_row = nullptr;
    float* ard_beta_full = nullptr;
    DevNetwork net;
    bool any_elementwise() const { return l2_w || group_id || l1_mask || ard_alpha; }
    void free_all() {
        cudaFree(l2_w); cudaFree(group_id); cudaFree(group_w); cudaFree(l1_mask); cudaFree(l1_w);
        cudaFree(ard_alpha); cudaFree(ard_beta_row); cudaFree(ard_beta_full);
        l2_w = group_w = l1_w = ard_alpha = ard_beta_row = ard_beta_full = nullptr;
        group_id = nullptr; l1_mask = nullptr;
        net.free_all();
    }
};

struct pmf_model_s {
    pmf_dims dims{};
    int M = 0, N = 0, K = 0, Kp = 0, lda = 0, Mp = 0, Np = 0;
    int n_sms = 148, cc_major = 0;
    std::string err;
    bool cuda_failed = false;
    bool have_data = false, have_noise = false;
    bool layout_set = false;                 // pmf_set_batch_layout has run (an identical layout is then a no-op)

    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    // data and factors
    float* A = nullptr;
    float *X = nullptr, *dX = nullptr, *accX = nullptr;
    float *Y = nullptr, *accY = nullptr;
    float *Xh = nullptr, *Xl = nullptr;   // TF32 operand split of X for the tcgen05 path (lazy)
    pmf::WideScratch wide{};              // K > 64 tensor-core path: operand scratch + the G' matrix (lazy)
    bool xsplit_valid = false;            // Xh / Xl match X (written by the update pass); else the launcher refreshes them
    bool grads_clean = false;             // dX, sg and the loss scalars were cleared by the previous epoch's update pass
    bool auto_tc = true;                     // PMF_KERNEL_AUTO picks the tcgen05 path when it applies
    // per-column noise description
    float* weight = nullptr;
    int32_t* colinfo = nullptr;
    int32_t* tc_cost_cum = nullptr;          // cumulative per-feature-tile cost of the tcgen05 data pass (n_jt + 1)
    float* thresholds = nullptr;
    float* acc_thr = nullptr;                // AdaGrad accumulators of the interior ordinal thresholds [PMF_MAX_RANGES][2]
    int n_ranges = 0;
    bool has_ordinal = false;
    // vector parameters  vp = [logsigma Np | mu Np | logdelta nbp | theta nbp]
    // shared gradients   sg = [dY Np*Kp | dlogsigma Np | dmu Np | dlogdelta nbp | dtheta nbp]
    float *vp = nullptr, *sg = nullptr, *accvp = nullptr, *regw = nullptr, *regc = nullptr;
    int64_t nbp = 0;
    int nb_max = 0;
    std::vector<BatchView> views;
    int32_t *bcol_off = nullptr, *bcol_view = nullptr, *bcol_nb = nullptr, *batch_of_sample = nullptr;
    // tcgen05 path with batch layers: sample orders / passes (see TcBatchDev); rebuilt lazily after the data,
    // the noise models or the batch layout changed
    std::vector<int32_t> bos_host;           // [n_views][M] batch ids as given
    std::vector<int32_t> tile_cost_host;     // per 128-feature tile
    pmf::TcBatchDev tcb{};
    bool tcb_valid = false;
    std::vector<void*> tcb_allocs;
    void free_tc_plan();
    int build_tc_plan();
    bool layer_reg_present[4] = {false, false, false, false};
    uint32_t frozen_layers = 0, frozen_regs = 0;
    SideReg reg[2];

    // Loss scalars and fit-control state.  `scalars` / `ctrl` point at the CURRENT buffers: the fused epoch loop of
    // pmf_fit rotates them through scalars_base (3 x SC_COUNT ring + 1 spare) and ctrl_base (2), everything else uses
    // the first of each.
    double* scalars_base = nullptr;
    pmf::FitControl* ctrl_base = nullptr;
    double* scalars = nullptr;
    pmf::FitControl* ctrl = nullptr;
    pmf::FitControl* ctrl_host = nullptr;   // pinned
    double* hist = nullptr;
    int hist_cap = 0;
    int cur_epoch = 1;
    int64_t launches = 0;
    float *col_ssq = nullptr, *col_cnt = nullptr, *col_sqerr = nullptr;
    int loss_grad_kernel = 0, loss_grad_precision = 0;
    bool profiling = false;
    std::vector<cudaEvent_t> prof_ev;    // pairs (start, stop) per bracketed data pass
    size_t prof_used = 0;

    // sample-sharded multi-GPU: NCCL communicator (opaque; loaded with dlopen) or null
    void* comm = nullptr;
    int comm_ranks = 1;

    size_t vp_len() const { return 2 * (size_t)Np + 2 * (size_t)nbp; }
    // shared gradients: [dY | dlogsigma | dmu | dlogdelta | dtheta | interior ordinal thresholds (2 per noise range)]
    size_t sg_len() const { return (size_t)Np * Kp + vp_len() + 2 * (size_t)pmf::PMF_MAX_RANGES; }
    float* g_thr() { return sg + (size_t)Np * Kp + vp_len(); }
    float* logsigma() { return vp; }
    float* mu() { return vp + Np; }
    float* logdelta() { return vp + 2 * (size_t)Np; }
    float* theta() { return vp + 2 * (size_t)Np + nbp; }
    float* g_Y() { return sg; }
    float* g_logsigma() { return sg + (size_t)Np * Kp; }
    float* g_mu() { return sg + (size_t)Np * Kp + Np; }
    float* g_logdelta() { return sg + (size_t)Np * Kp + 2 * (size_t)Np; }
    float* g_theta() { return sg + (size_t)Np * Kp + 2 * (size_t)Np + nbp; }

    int realloc_vectors(int new_nbp);
    int run_data_pass(pmf::DataPassParams& p, int kind, int precision);
    void fill_factor_params(int which, pmf::FactorUpdateParams& q);
    int run_network_reg(int which, const int* stop);
    int run_reg_multi(bool x_side, bool y_side, bool vectors, const int* stop);
    int run_update_multi(bool upd_X, bool upd_Y, bool upd_layers, bool upd_noise, float lr, float eps, const int* stop);
    // the fused epoch pass of pmf_fit (termination test + penalties + update in one launch); values_only: just the
    // penalty values at the current parameters into `scalars` (first epoch of a fit)
    void fill_epoch_pass(pmf::MultiPassParams& mp, const pmf_fit_opts* o, bool values_only);
    int run_penalty_values(const pmf_fit_opts* o);
    int run_fused_epoch(const pmf_fit_opts* o, const pmf::FusedControl& fc);
    int exchange_gradients();   // all-reduce of sg and of the rank-local loss scalars on `stream`
};
