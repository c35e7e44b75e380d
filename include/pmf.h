/*
 * libpmf -- C ABI of the B200-native PathMatFac fit-loop hot path.
 *
 * Drop-in boundary: the reference has no FFI; its boundary is the Julia call
 *     h = MF.fit!(model.matfac, model.data; update_X, update_Y, update_col_layers, ...,
 *                 opt, max_epochs, epoch, rel_tol, abs_tol, ...)      (src/fit.jl:24-36)
 * wrapped by mf_fit! (src/fit.jl:9-38) and driven by mf_fit_adapt_lr! (src/fit.jl:46-75).
 * A Julia shim (julia/PathMatFacB200.jl) marshals a PathMatFacModel
 * into the calls below through `ccall`; the Python host mirror in pathmatfac.jl_b200/
 * binds the same symbols through ctypes.  Plain pointers and sizes only.
 *
 * Conventions
 *  - Every function returns 0 on success, a negative pmf_status otherwise; the message is
 *    available from pmf_last_error(h) (h may be NULL for errors of pmf_create).
 *  - Host arrays are in the reference's (Julia, column-major) layout: X is K x M, Y is K x N,
 *    data is M x N with NaN = missing, batch values are n_b x N_v.  Index vectors are int32
 *    and 0-based; column ranges are [start, stop) 0-based.
 *  - The caller owns every host buffer; the library copies at set_* and copies back at
 *    get_*; no pointer is retained after a call returns.  Device memory belongs to the handle.
 *  - Not thread-safe on one handle.  There is no CPU execution path: every entry point that
 *    computes needs a CUDA device (sm_100a) and fails with PMF_ERR_CUDA otherwise.
 */
#ifndef PMF_H
#define PMF_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pmf_model_s* pmf_handle;

typedef enum {
    PMF_OK = 0,
    PMF_ERR_ARG = -1,      /* bad argument / inconsistent shapes            */
    PMF_ERR_CUDA = -2,     /* CUDA runtime error (sticky on the handle)     */
    PMF_ERR_STATE = -3,    /* call order (e.g. fit before set_data)         */
    PMF_ERR_ALLOC = -4
} pmf_status;

/* Noise-model ids: VALID_LOSSES, src/util.jl:128. */
typedef enum {
    PMF_NORMAL = 0, PMF_BERNOULLI = 1, PMF_POISSON = 2, PMF_ORDINAL3 = 3,
    PMF_BERNOULLI_SQ_HINGE = 4, PMF_ORDINAL_SQ_HINGE3 = 5
} pmf_dist;

/* h["term_code"] of MF.fit! (only "loss_increase" is consumed, src/fit.jl:63). */
typedef enum {
    PMF_TERM_MAX_EPOCHS = 0, PMF_TERM_ABS_TOL = 1, PMF_TERM_REL_TOL = 2,
    PMF_TERM_LOSS_INCREASE = 3, PMF_TERM_NONFINITE = 4
} pmf_term_code;

/* Which hand-written kernel computes the data pass. */
typedef enum {
    PMF_KERNEL_AUTO = 0,   /* tcgen05 kernels when the shape supports them AND the problem is large enough
                              for their TF32 gradient contractions to stay within 1e-4 (pmf_fit_opts.precision),
                              else FFMA */
    PMF_KERNEL_FFMA = 1,   /* FP32 CUDA-core tile kernel (exact FP32 products)   */
    PMF_KERNEL_TC = 2      /* tcgen05/TMEM kernels (split-BF16 Z, TF32 gradients); an error on shapes they do
                              not support, never a silent fallback */
} pmf_kernel_kind;

typedef struct {
    int32_t M;        /* samples owned by this handle (rows of the data, columns of X) */
    int32_t N;        /* features                                                      */
    int32_t K;        /* latent dimension                                              */
    int32_t device;   /* CUDA device ordinal                                           */
} pmf_dims;

/* size(model.data), size(model.matfac.X,1)  (src/model.jl:46, src/fit.jl:142). */
int pmf_create(const pmf_dims* dims, pmf_handle* out);
int pmf_destroy(pmf_handle h);
/* mf_fit! on a host-resident model creates and destroys a handle per call (src/fit.jl:9-38 with the model moved
 * by gpu()/cpu() around it), and cudaMalloc / cudaFree would otherwise dominate short fits.  A destroyed handle
 * therefore parks its data buffer (>= 64 MB; at most two are kept per process) for the next pmf_create of the same
 * size on the same device, and every smaller device block the library frees is kept in a size-keyed cache (at most
 * 1 GB per process; handed out zero-filled; PMF_ALLOC_CACHE=0 in the environment disables it).  This call returns
 * all of that memory to the driver. */
int pmf_release_cached_memory(void);

/* Host-side preview of how the tcgen05 data pass lays out a model with batch layers (no device work, no handle;
 * the same planner every handle runs).  BatchArray (src/batch_array.jl:5-15, 47-79) puts no order on the
 * samples of a batch; the kernel wants every 16-sample chunk inside one batch, so a view that does not come
 * that way gets a SAMPLE ORDER (samples stably sorted by batch id, every batch padded to a multiple of 16
 * positions; order 0 = identity), and every 128-column tile is walked once per order its views need (a PASS).
 *   bos:         [n_views][M] batch of each sample, 0-based
 *   view_order:  [n_views] order of each view                       (may be NULL)
 *   perm:        [n_orders][n_pos] position -> sample, -1 = padding (may be NULL; capacity in elements)
 *   pass_feat0 / pass_order: [n_pass] first column and order of each pass (may be NULL)
 *   chunk_batch: [n_views][n_pos / 16] batch of each chunk in the view's own order (may be NULL) */
int pmf_plan_batch_orders(int32_t M, int32_t N, int32_t n_views, const int32_t* col_start, const int32_t* col_stop,
                          const int32_t* n_batches, const int32_t* bos, int32_t* n_orders, int32_t* n_pass,
                          int32_t* n_pos, int32_t* view_order, int32_t* perm, int64_t perm_cap, int32_t* pass_feat0,
                          int32_t* pass_order, int32_t pass_cap, uint16_t* chunk_batch, int64_t chunk_cap);
const char* pmf_last_error(pmf_handle h);
/* Library / build identification ("libpmf <ver> sm_100a"). */
const char* pmf_version(void);

/* Run all of the handle's work on an externally owned CUDA stream (cudaStream_t as void*),
 * e.g. torch's current stream so that a torch.distributed collective on the gradient buffer
 * is ordered with the kernels.  NULL = the handle's own stream. */
int pmf_set_stream(pmf_handle h, void* cuda_stream);

/* model.data (M x N column-major, NaN = missing)                       src/model.jl:54 */
int pmf_set_data(pmf_handle h, const float* A_host);
/* model.matfac.X (K x M), model.matfac.Y (K x N)                  src/fit.jl:133-136 */
int pmf_set_factors(pmf_handle h, const float* X_host, const float* Y_host);
int pmf_get_factors(pmf_handle h, float* X_host, float* Y_host);

/* model.matfac.noise_model: CompositeNoise col_ranges / noises / weights / ext_thresholds
 * (src/fit.jl:224-245, src/regularizers.jl:756-757, src/impute.jl:15-35).
 * thresholds: 4 floats per range ([-Inf, t1, t2, +Inf] for the 3-level ordinal types;
 * ignored otherwise).  weight: N per-column loss weights (MF.set_weight!, src/fit.jl:157). */
int pmf_set_noise(pmf_handle h, int32_t n_ranges, const int32_t* col_start,
                  const int32_t* col_stop, const int32_t* dist_code,
                  const float* thresholds, const float* weight);

/* col_transform.layers[1].logsigma, .layers[3].mu          src/layers.jl:9-90 */
int pmf_set_col_params(pmf_handle h, const float* logsigma, const float* mu);
int pmf_get_col_params(pmf_handle h, float* logsigma, float* mu);

/* BatchScale.logdelta / BatchShift.theta share one BatchArray layout
 * (src/batch_array.jl:5-15, src/layers.jl:95-214).  Declare all batched views first
 * (ascending, non-overlapping column ranges), then set values per view.
 * batch_of_sample: n_views x M (row v = ordinal, 0-based, of each sample's batch in view v,
 * in `unique()` first-appearance order = column of its nonzero in row_batches[v]). */
int pmf_set_batch_layout(pmf_handle h, int32_t n_views, const int32_t* col_start,
                         const int32_t* col_stop, const int32_t* n_batches,
                         const int32_t* batch_of_sample);
/* values: n_b x N_v column-major (either pointer may be NULL to leave it unchanged). */
int pmf_set_batch_values(pmf_handle h, int32_t view, const float* logdelta, const float* theta);
int pmf_get_batch_values(pmf_handle h, int32_t view, float* logdelta, float* theta);

/* FrozenLayer / FrozenRegularizer per slot (1 ColScale, 2 BatchScale, 3 ColShift,
 * 4 BatchShift -> bits 0..3)          src/layers.jl:299-363, src/regularizers.jl:950-1004 */
int pmf_set_frozen(pmf_handle h, uint32_t frozen_layer_mask, uint32_t frozen_reg_mask);

/* ---- regularisers ---------------------------------------------------------------------
 * `which`: 0 = X_reg (rows are samples), 1 = Y_reg (rows are features).  Mixture weights of
 * CompositeRegularizer (src/regularizers.jl:616-649) are passed as `p`; a bare regulariser
 * uses p = 1.  pmf_clear_reg removes every component of that side (the `x->0` closures). */
int pmf_clear_reg(pmf_handle h, int32_t which);
/* L2Regularizer: 0.5 sum_k w_k sum_i P[k,i]^2                  src/regularizers.jl:11-55 */
int pmf_set_reg_l2(pmf_handle h, int32_t which, const float* w_K, float p);
/* GroupRegularizer: contiguous ranges, per-group K-vectors (n_groups x K row-major)
 *                                                             src/regularizers.jl:345-456 */
int pmf_set_reg_group(pmf_handle h, int32_t which, int32_t n_groups, const int32_t* start,
                      const int32_t* stop, const float* w_groups_K, float p);
/* SelectiveL1Reg: l1_idx is K x n column-major bytes          src/regularizers.jl:106-163 */
int pmf_set_reg_sel_l1(pmf_handle h, int32_t which, const uint8_t* l1_idx, const float* w_K,
                       float p);
/* ARDRegularizer: per-range alpha, beta                       src/regularizers.jl:526-609 */
int pmf_set_reg_ard(pmf_handle h, int32_t which, int32_t n_ranges, const int32_t* start,
                    const int32_t* stop, const float* alpha, const float* beta);
/* FeatureSetARDReg value/pullback: alpha (n), beta (K x n col-major)
 *                                                            src/featureset_ard.jl:135-150 */
int pmf_set_reg_fsard(pmf_handle h, int32_t which, const float* alpha, const float* beta);
/* NetworkRegularizer (src/regularizers.jl:169-338): for each factor k the symmetric blocks
 * AA_k (n x n), AB_k (n x nv_k), BB_k (nv_k x nv_k) in CSR (== the reference's CSC because the
 * Laplacian is symmetric; AB is passed as CSR of AB_k, i.e. n rows).  The K matrices are
 * concatenated: *_rowptr has one (rows+1)-long segment per factor with nnz offsets local to
 * the factor; nv[k] = number of virtual nodes.  x_virtual: concatenated warm starts
 * (sum nv) -- the sign-flipped vector the reference stores (may be NULL = zeros).
 * cg_rtol / cg_atol <= 0 select sqrt(eps(Float32)) (Krylov.jl default); cg_itmax <= 0 -> 2 nv. */
int pmf_set_reg_network(pmf_handle h, int32_t which, const int32_t* nv,
                        const int32_t* aa_rowptr, const int32_t* aa_col, const float* aa_val,
                        const int32_t* ab_rowptr, const int32_t* ab_col, const float* ab_val,
                        const int32_t* bb_rowptr, const int32_t* bb_col, const float* bb_val,
                        const float* x_virtual, float p, float cg_rtol, float cg_atol,
                        int32_t cg_itmax);
int pmf_get_network_virtual(pmf_handle h, int32_t which, float* x_virtual);

/* Layer regularisers (SequenceReg, src/regularizers.jl:896-938): ColParamReg for slots 1/3
 * (per-column weight and centre, expanded from the per-view scalars) and BatchArrayReg for
 * slots 2/4 (per-view, per-batch weight and centre; n_b values per view, concatenated).
 * slot is 1..4; NULL weight removes the regulariser (`x->0`). */
int pmf_set_layer_reg_col(pmf_handle h, int32_t slot, const float* weight_N, const float* center_N);
int pmf_set_layer_reg_batch(pmf_handle h, int32_t slot, const float* weight_b, const float* center_b);

/* Flux AdaGrad state (opt.acc, src/optimizers.jl:6-13): which = 0 X, 1 Y, 2 logsigma, 3 mu,
 * 4 logdelta(view), 5 theta(view).  Same host layouts as the parameters.  pmf_reset_opt_state
 * re-initialises every accumulator to epsilon (a fresh `AdaGrad(lr)`, src/fit.jl:41-43). */
int pmf_reset_opt_state(pmf_handle h, float epsilon);
int pmf_get_opt_state(pmf_handle h, int32_t which, int32_t view, float* acc_host);
int pmf_set_opt_state(pmf_handle h, int32_t which, int32_t view, const float* acc_host);

typedef struct {
    double data, x_reg, y_reg, layer_reg, total;
} pmf_losses;

/* Parity hook: one loss+gradient evaluation at the current parameters, no update.
 * Any gradient pointer may be NULL.  Gradients include the regularisers' pullbacks when
 * `include_reg` != 0.  dX: K x M, dY: K x N, dlogsigma/dmu: N.  Batch gradients are read
 * with pmf_get_batch_grads afterwards.                    (MF.likelihood + Zygote, App. B) */
int pmf_loss_grad(pmf_handle h, int32_t include_reg, pmf_losses* out, float* dX, float* dY,
                  float* dlogsigma, float* dmu);
int pmf_get_batch_grads(pmf_handle h, int32_t view, float* dlogdelta, float* dtheta);
/* d loss / d(t1, t2) of every noise range after pmf_loss_grad ([n_ranges][2]; zero for non-ordinal ranges) and the
 * current extended thresholds [n_ranges][4] (after a fit with update_noise_models: the trained values, which the
 * caller writes back into noise.ext_thresholds; src/fit.jl:228-242, src/impute.jl:15-24). */
int pmf_get_threshold_grads(pmf_handle h, int32_t n_ranges, float* dthr);
int pmf_get_thresholds(pmf_handle h, int32_t n_ranges, float* thresholds);

typedef struct {
    /* kwargs of mf_fit! / MF.fit! (src/fit.jl:9-36, 923-939) */
    int32_t max_epochs;          /* last epoch index allowed (inclusive)               */
    int32_t epoch;               /* first epoch index (1-based; resume point)          */
    float   lr;                  /* opt.eta                                            */
    float   adagrad_eps;         /* opt.epsilon (1e-8)                                 */
    double  rel_tol, abs_tol;
    int32_t update_X, update_Y, update_col_layers;
    int32_t kernel;              /* pmf_kernel_kind                                    */
    int32_t precision;           /* tensor-core kernels only (PMF_KERNEL_FFMA is exact FP32 throughout):
                                    0, 1 (identical): Z = X'Y at FP32 level -- a two-term BF16 split of both
                                       operands (v = h + l, h = bf16(v), l = bf16(v - h)) and the three
                                       contractions Yh Xh + Yl Xh + Yh Xl with FP32 accumulation (dropped
                                       term: 2^-17 of a product), for K <= 64 and K > 64 alike -- so loss and
                                       column / batch gradients agree with FP32 to ~1e-6; the contractions
                                       dX = Y G', dY = X G are single-pass TF32 with round-to-nearest
                                       operands (FP32 accumulation): ~1-2e-4 relative on small problems,
                                       3-4e-5 at 10 000 x 30 000.  PMF_KERNEL_AUTO therefore picks these
                                       kernels only from M N >= 6e6 (K/64)^2, min(M, N) >= 2000, where the
                                       measured gradient error is below 1e-4.
                                    2: experiments -- only the leading term Yh Xh of Z (K <= 64; ignored
                                       for K > 64)                                                       */
    int32_t check_every;         /* epochs launched between host checks of the stop flag */
    int32_t no_terminate;        /* bench hook: run every epoch up to max_epochs, never stop on
                                    a tolerance / loss-increase test (losses still recorded) */
    int32_t update_noise_models; /* src/fit.jl:14 (true in every call of the reference): train the interior
                                    thresholds of the ordinal noise models with the same optimiser
                                    (the trainable noise parameters are INFERRED); no effect
                                    on models without ordinal columns.  Default 1.                     */
    int32_t alternating;         /* SURVEY App. D1, the unknown epoch order of MF.fit!: 0 (default) = one pass,
                                    simultaneous step of every parameter; 1 = column-side step (Y, layers,
                                    thresholds) from the first pass, then the row-side step (X) from a second
                                    pass at the new column-side parameters (pmf_fit only)              */
} pmf_fit_opts;

typedef struct {
    int32_t term_code;           /* pmf_term_code                                      */
    int32_t epochs;              /* h["epochs"]: index of the last epoch evaluated     */
    int32_t n_recorded;          /* number of per-epoch records written below          */
    int32_t capacity;            /* in: length of the arrays below (may be 0)          */
    double* loss_total;          /* per-epoch losses (host, caller-owned, may be NULL) */
    double* loss_data;
    double* loss_x_reg;
    double* loss_y_reg;
    double* loss_layer_reg;
    float   device_ms;           /* CUDA-event time of the epochs on the handle's stream */
    int64_t kernel_launches;     /* kernels launched by this call                       */
} pmf_history;

void pmf_default_fit_opts(pmf_fit_opts* o);
/* One MF.fit! call: epochs `epoch`..`max_epochs` of loss+gradient+AdaGrad update with the
 * termination test evaluated on the device.                             src/fit.jl:24-36 */
int pmf_fit(pmf_handle h, const pmf_fit_opts* opts, pmf_history* out);

/* Multi-rank (sample-sharded) epoch: begin = zero gradients + data pass + X-side penalties on
 * this rank's samples; the caller then sums the shared gradient buffer and the loss scalars
 * over ranks (NCCL all-reduce on the pointers below); end = Y-side / layer penalties,
 * termination test and the AdaGrad step.  pmf_fit_poll copies out the stop flag / history. */
int pmf_epoch_begin(pmf_handle h, const pmf_fit_opts* opts);
int pmf_epoch_end(pmf_handle h, const pmf_fit_opts* opts);
int pmf_fit_start(pmf_handle h, const pmf_fit_opts* opts);
int pmf_fit_poll(pmf_handle h, pmf_history* out, int32_t* stopped);
/* Device pointers (on the handle's device) of the buffers a multi-rank caller all-reduces:
 * grads = [dY (Np*Kp) | dlogsigma | dmu | dlogdelta | dtheta] floats, scalars = doubles. */
int pmf_shared_grad_buffer(pmf_handle h, void** dev_ptr, int64_t* n_floats);
int pmf_shared_scalar_buffer(pmf_handle h, void** dev_ptr, int64_t* n_doubles);

/* FeatureSetARD outer step (update_A!, src/featureset_ard.jl:214-294) for one view:
 * S is L x N_v CSR (rows = feature sets); A and ssq_grad (the ISTAOptimiser state) are passed
 * as K x L column-major buffers (the reference's L x K matrices transposed; A is out only --
 * it is zeroed first like the reference, featureset_ard.jl:286); lambda has K entries.  Uses the
 * handle's FSARD alpha (pmf_set_reg_fsard) and the handle's
 * current Y[:, col_start:col_stop]; writes beta = beta0 (v0 + A'S) into the handle's FSARD
 * beta for those columns and returns the best loss and the epochs run. */
int pmf_fsard_update_A(pmf_handle h, int32_t col_start, int32_t col_stop, int32_t L,
                       const int32_t* S_rowptr, const int32_t* S_col, const float* S_val,
                       float* A_host, float* ssq_grad_host, const float* lambda_K,
                       float lr, float alpha0, float v0, int32_t max_epochs, int32_t term_iter,
                       float atol, double* best_loss, int32_t* epochs_run);
int pmf_get_fsard_beta(pmf_handle h, float* beta_host /* K x N */);

/* ---- sample-sharded multi-GPU (one process per GPU) -----------------------------
 * The path has one exchange step per epoch: the sum over ranks of the shared gradient buffer
 * [dY | dlogsigma | dmu | dlogdelta | dtheta] and of the rank-local loss scalars.  With a
 * communicator attached, pmf_fit issues it itself (ncclAllReduce on the handle's stream, between
 * the data pass and the update), so the epoch loop never returns to the host.  NCCL is loaded
 * with dlopen("libnccl.so.2") on first use; single-GPU callers do not need it.
 * pmf_comm_unique_id: rank 0 creates the 128-byte ncclUniqueId, the host language broadcasts it
 * (MPI / torch.distributed / Distributed.jl), every rank calls pmf_comm_init_rank. */
int pmf_comm_unique_id(uint8_t id_out[128]);
int pmf_comm_init_rank(pmf_handle h, int32_t n_ranks, int32_t rank, const uint8_t id[128]);
int pmf_comm_destroy(pmf_handle h);

/* Next-tier O(MN) passes of the staging code that reuse the tile kernel:
 * per-column sum_i (dl/dz)^2 (MF.batched_column_ssq_grads, src/fit.jl:166) and per-column
 * count of finite entries (MF.column_nonnan, src/fit.jl:140). */
int pmf_column_stats(pmf_handle h, float* ssq_grads_N, float* nonnan_N);
/* MF.link_col_sqerr (src/fit.jl:138-139,444-446): per column sum_i (D_ij - forward_ij)^2 over finite
 * entries, forward = the full layer stack at the current parameters, and MF.column_nonnan (:140,447).
 * [The link of the non-normal noise models lives in MatFac.jl; identity is used for every column.] */
int pmf_link_col_sqerr(pmf_handle h, float* sqerr_N, float* nonnan_N);
/* ba_map(d -> isfinite.(d), theta, data) and ba_map(MF.sqerr_func, theta, model, data)
 * (src/batch_array.jl:320-334; src/fit.jl:332,355-356,454-456): per batched view v the n_b x N_v
 * column-major tables of segmented column sums (count of finite entries; squared error in link space).
 * count / sqerr: arrays of n_views host pointers (an entry or the array itself may be NULL). */
int pmf_batch_stats(pmf_handle h, int32_t n_views, float* const* count, float* const* sqerr);

/* Test / bench hooks (no reference counterpart). */
/* kernel kind and precision used by pmf_loss_grad */
int pmf_set_loss_grad_kernel(pmf_handle h, int32_t kernel, int32_t precision);
/* enable != 0: bracket every data-pass launch of pmf_fit with CUDA events on the handle's
 * stream; pmf_get_profile returns the number of bracketed launches since enabling and their
 * mean / min duration in milliseconds. */
int pmf_set_profiling(pmf_handle h, int32_t enable);
int pmf_get_profile(pmf_handle h, int32_t* n_launches, float* mean_ms, float* min_ms);
/* With PMF_GUARD=1 in the environment (read at the first allocation) every device buffer of the library carries a
 * 1 KiB guard zone on each side; this verifies them all: number of live buffers and of corrupted guard bytes.
 * Returns PMF_ERR_STATE when guards are not enabled. */
int pmf_check_guards(int64_t* n_buffers, int64_t* n_corrupt_bytes);

#ifdef __cplusplus
}
#endif
#endif /* PMF_H */
