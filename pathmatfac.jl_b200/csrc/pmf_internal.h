// Internal device-side views shared by the kernels and the C-ABI translation unit.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pmf {

// guard.cu: every device allocation of the library goes through these (a size-keyed cache of freed blocks; guard
// zones instead when PMF_GUARD=1).  The macros below route the remaining direct calls of the translation units that
// include this header.
cudaError_t guarded_malloc(void** p, size_t bytes);
cudaError_t guarded_free(void* p);
void release_alloc_cache();      // every parked block back to the driver (pmf_release_cached_memory)

// Number of double scalars in the shared scalar buffer.
enum { SC_DATA = 0, SC_XREG = 1, SC_YREG = 2, SC_LAYERREG = 3, SC_COUNT = 8 };

// Everything the fused data pass reads / accumulates.  Layouts:
//   A  [N][lda]   feature-major, lda = roundup(M,32), padding = NaN (missing)
//   X  [Mp][Kp], Y [Np][Kp]   row = sample / feature, k contiguous, zero padded
//   per-column vectors have Np entries; batch tables are [view columns][n_b] flattened.
struct DataPassParams {
    int M, N, Kp, lda, Mp, Np;
    const float* __restrict__ A;
    const float* __restrict__ X;
    const float* __restrict__ Y;
    const float* __restrict__ logsigma;
    const float* __restrict__ mu;
    const float* __restrict__ weight;
    const int32_t* __restrict__ colinfo;      // dist | range_id << 8
    const float* __restrict__ thresholds;     // [n_ranges][4]
    const int32_t* __restrict__ tc_cost_cum;  // [ceil(N/128) + 1] cumulative tile cost per feature tile (tcgen05 path)
    // batch layers (null / -1 when absent)
    int n_batch_views;
    int nb_max;                               // max n_batches over views
    const int32_t* __restrict__ bcol_off;     // [Np] offset of column's n_b block, -1 = unbatched
    const int32_t* __restrict__ bcol_view;    // [Np] view index of the column (-1)
    const int32_t* __restrict__ bcol_nb;      // [Np] n_batches of the column's view
    const int32_t* __restrict__ batch_of_sample;  // [n_views][Mp]
    const float* __restrict__ logdelta;
    const float* __restrict__ theta;
    // outputs (accumulated with atomics; zeroed by the caller)
    float* dX;          // [Mp][Kp]
    float* dY;          // [Np][Kp]
    float* dlogsigma;   // [Np]
    float* dmu;         // [Np]
    float* dlogdelta;   // batch table layout
    float* dtheta;
    float* dthr;        // [n_ranges][2] gradients of the interior ordinal thresholds, or null (thresholds not trained)
    double* scalars;    // SC_* ; data loss accumulated into scalars[SC_DATA]
    // optional per-column statistics pass (pmf_column_stats)
    float* col_ssq;     // [Np] sum_i g^2 or null
    float* col_cnt;     // [Np] count of finite entries or null
    float* col_sqerr;   // [Np] sum_i (a - z4)^2 over finite entries or null (statistics pass only)
    const int* stop_flag;   // device flag: non-zero => the fit has terminated, do nothing
    int sample_chunks;      // grid.y: number of sample chunks per feature tile
    float ordinal_eps, hinge_margin;
};

// Elementwise regulariser + AdaGrad pass over a factor matrix P [n_pad][Kp].
struct FactorUpdateParams {
    int n, Kp, K;
    float* P;
    const float* grad;          // data gradient (already reduced over ranks)
    float* grad_out;            // if non-null: data+reg gradient is written here (parity hook)
    float* acc;                 // AdaGrad accumulator
    // quadratic penalties
    const float* l2_w;          // [Kp] (mixture weight folded in) or null
    const int32_t* group_id;    // [n] or null
    const float* group_w;       // [n_groups][Kp]
    // selective L1
    const uint8_t* l1_mask;     // [n][Kp] or null
    const float* l1_w;          // [Kp]
    // (FS)ARD
    const float* ard_alpha;     // [n] or null
    const float* ard_beta_row;  // [n] or null
    const float* ard_beta_full; // [n][Kp] or null
    double* loss_out;           // scalar accumulated (pre-update value of the penalty)
    int do_update;
    int loss_after;             // fused epoch pass: the penalty VALUE is taken at the updated parameters (next epoch's loss)
    float lr, eps;
    const int* stop_flag;
    // optional by-products of the update pass
    float* Ph;                  // rna_tf32(P) and P - rna_tf32(P) of the UPDATED parameters (operands of the
    float* Pl;                  //   tcgen05 data pass) or null
    float* zero_buf;            // gradient buffer to clear once it has been consumed ([n_pad][Kp]) or null
    int n_pad;
};

constexpr int PMF_MAX_RANGES = 64;     // noise ranges whose thresholds can be trained (tail of the shared gradient buffer)

// 1-D parameters (logsigma | mu | logdelta | theta) with optional quadratic penalty.
struct VectorUpdateParams {
    int n;
    float* p;
    const float* grad;
    float* grad_out;
    float* acc;
    const float* reg_w;         // per-element weight or null
    const float* reg_c;         // per-element centre
    double* loss_out;
    int reg_active;             // 0 when the slot's regulariser is frozen / absent
    int do_update;
    int loss_after;             // as in FactorUpdateParams
    float lr, eps;
    const int* stop_flag;
    float* zero_buf;            // gradient segment to clear once consumed, or null
};

// Several factor / vector passes in ONE launch (the per-epoch penalty pass and the per-epoch update
// pass): segments share the grid in proportion to their size, so the small column-parameter work
// hides inside the factor passes instead of paying a launch each.
struct MultiPassParams {
    int nf, nv;
    FactorUpdateParams f[2];
    VectorUpdateParams v[4];
    double* zero_scalars;       // SC_COUNT doubles cleared by block 0 after the pass (next epoch's accumulators), or null
    // interior ordinal thresholds (update pass only): th [n_ranges][4], gradients / accumulators [n_ranges][2]
    float* thr;
    float* thr_grad;
    float* thr_acc;
    int thr_ranges;             // 0: nothing to do
    int thr_update;             // AdaGrad step (else only the gradients are cleared)
    float thr_lr, thr_eps;
};

// Fused epoch pass (pmf_fit): termination test + penalties + AdaGrad step in ONE launch.  Every block evaluates the
// test from the epoch's loss scalars and the control state of the PREVIOUS epoch (read-only during the launch: state
// and scalars are ring buffers), block 0 writes the next state and the history record.
struct FitControl;
struct FusedControl {
    const FitControl* cin;    // state left by the previous epoch
    FitControl* cout;         // state for the next epoch
    const double* sc;         // this epoch's loss scalars (data loss from the data pass, penalties from the previous
                              //   epoch's pass, which evaluated them at the updated parameters)
    double* sc_zero;          // the scalars of the epoch after next: cleared here
    double* hist;
    int hist_cap, epoch, max_epochs;
    double rel_tol, abs_tol;
};

struct FitControl {
    int stop;            // set by the control kernel when a termination test fires
    int term_code;
    int epochs;          // index of the last epoch evaluated
    int n_recorded;
    double prev_loss;
    int have_prev;
    int pad;
};

struct CsrBlock {        // K concatenated CSR matrices
    const int32_t* rowptr;   // per factor segment of (rows+1) entries, local offsets
    const int32_t* col;
    const float* val;
    const int64_t* rowptr_base;  // [K] start of factor's rowptr segment
    const int64_t* nnz_base;     // [K] start of factor's col/val segment
};

struct NetworkParams {
    int n, Kp, K;
    const float* P;          // [n_pad][Kp]
    float* grad;             // += p * (AA y + AB u)
    CsrBlock AA, AB, BB, ABt;  // ABt: CSR of AB^T (nv rows)
    const int32_t* nv;       // [K]
    const int64_t* virt_base;    // [K] offset into the virtual-node vectors
    float* yt;               // [K][ld] transposed copy of P (scratch)
    float* gt;               // [K][ld] AA y + AB u per factor (scratch)
    int ld;
    float* u;                // x_virtual (sign-flipped convention of the reference)
    float* work;             // 4 * sum(nv) scratch: r, pvec, Ap, rhs
    int64_t nv_total;
    float p;
    float rtol, atol;
    int itmax;
    double* loss_out;
    const int* stop_flag;
};

// launchers (each defined next to its kernel)
cudaError_t launch_data_pass_ffma(const DataPassParams& p, cudaStream_t s, int n_sms);
// tcgen05 path (fused_tc.cu): K <= 64 (operands zero padded to 64).  Xh = rna_tf32(X) ([Mp][64] FP32) and
// Xl = bf16([Xh | X - Xh]) ([Mp][128] BF16) are scratch operands kept by the update pass (refreshed by the
// launcher when `refresh_split`).  precision: 0/1 = TF32 + BF16 first-order corrections for Z and TF32
// gradients, 2 = TF32 everywhere.
//
// Batch layers: the epilogue keeps the parameters of ONE batch per thread (column) in registers and switches
// them between 16-sample chunks, so it wants every chunk to lie in one batch.  A view whose batch ids do not
// come that way gets its own SAMPLE ORDER: samples stably sorted by batch id, every batch padded to a multiple
// of 16 positions (padding positions hold no sample: A is NaN there).  Order 0 is the identity.  A PASS is a
// (128-feature tile, order) pair: a tile whose columns need several orders is walked once per order, over a
// copy of A (A_tc) whose rows are laid out in the pass' order and are NaN for the columns another pass serves.
// Per order there is a copy of the X operands and of the dX accumulator; after the data pass the dX copies are
// gathered back (combine kernel).  When every view is chunk-uniform as given (`direct`) there is one order and
// none of the copies exist.
struct TcBatchDev {
    int n_orders, n_pass, n_views;
    int n_pos;                    // sample positions per order (multiple of 128; = Mp when direct)
    int n_used;                   // positions in use: max over the orders (= M when direct)
    bool direct;
    const float* A_tc;            // [n_pass * 128][n_pos]                     (null when direct)
    const int32_t* perm;          // [n_orders][n_pos] position -> sample, -1  (null when direct)
    const int32_t* pos;           // [n_orders][M]     sample -> position      (null when direct)
    float* Xh;                    // [n_orders * n_pos][64]                    (null when direct)
    float* Xb;                    // [n_orders * n_pos][128] BF16              (null when direct)
    float* dX;                    // [n_orders * n_pos][Kp], zero between passes (null when direct)
    const int32_t* pass_feat0;    // [n_pass]
    const int32_t* pass_order;    // [n_pass]
    const int32_t* view_order;    // [n_views]
    const int32_t* cost_cum;      // [n_pass + 1]
    const uint16_t* boc;          // [n_views][n_pos/16] batch of a chunk in the view's order
};
// 64 < K <= 256 (wide_tc.cu): operand scratch of the three-kernel tensor-core pass.  Kq = roundup(Kp, 32).
struct WideScratch {
    float* Xh;     // [Mp][Kq]   rna_tf32(X)
    void* Xb;      // [Mp][2 Kq] BF16, per 32-factor slab [Xh | Xl]
    float* Yh;     // [Np][Kq]
    void* Yb;      // [Np][2 Kq] BF16, per slab [Yl | Yh]
    float* G;      // [N][lda]   w_j sigma_j dloss/dz
};
bool wide_supported(const DataPassParams& p);
size_t wide_scratch_floats(int rows_pad, int Kp);      // floats of Xh / Xb (each) for `rows_pad` rows
cudaError_t launch_data_pass_wide(const DataPassParams& p, const WideScratch& ws, int precision, cudaStream_t s, int n_sms,
                                  int* n_launches);
bool tc_supported(const DataPassParams& p);
// `bp`: null for a model without batch layers.  *n_launches receives the number of kernels launched.
cudaError_t launch_data_pass_tc(const DataPassParams& p, float* Xh, float* Xl, bool refresh_split, int precision,
                                cudaStream_t s, int n_sms, const TcBatchDev* bp, int* n_launches);
// A_tc rows for the plan (fused_tc.cu)
cudaError_t launch_build_a_tc(const DataPassParams& p, const TcBatchDev& bp, float* A_tc, cudaStream_t s);
cudaError_t launch_multi_pass(const MultiPassParams& p, cudaStream_t s, int n_sms);
cudaError_t launch_fused_epoch_pass(const MultiPassParams& p, const FusedControl& fc, cudaStream_t s, int n_sms);
cudaError_t launch_control(FitControl* ctrl, const double* scalars, double* hist, int hist_cap,
                           int epoch, int max_epochs, double rel_tol, double abs_tol, cudaStream_t s);
// four launches (transpose in, virtual-node solves, rows, transpose-add); nv_max = max virtual nodes over the factors
cudaError_t launch_network_reg(const NetworkParams& p, cudaStream_t s, int n_sms, int nv_max);

}  // namespace pmf

#ifndef PMF_NO_GUARD_MACROS
#define cudaMalloc(p, n) ::pmf::guarded_malloc(reinterpret_cast<void**>(p), (n))
#define cudaFree(p) ::pmf::guarded_free((void*)(p))
#endif
