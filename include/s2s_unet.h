/*
 * s2s_unet.h — C ABI of the B200-native U-Net hot path (libs2s_unet.so).
 *
 * This is the drop-in boundary for the U-Net fit / predict / skill path behind the
 * reference's tune_*.py scripts.  The reference (emileDesmaili/s2s-ismr-unet) has no
 * FFI of its own: its seam is the Keras object API used by utils/training.py.  Every
 * entry point below cites the reference call it replaces (file:line into the reference).
 * The Python host layer (s2s-ismr-unet_b200/) binds these symbols with ctypes and
 * re-exposes the reference's own class / function names on top.
 *
 * Conventions
 *   - plain C types only; no C++ / torch types cross the boundary;
 *   - every function returns 0 on success or a negative s2s_status; the message is
 *     available from s2s_last_error() (thread-local);
 *   - pointers named *_dev are DEVICE pointers owned by the caller unless documented
 *     as "borrowed" (handle-owned arenas handed out for inspection / NCCL);
 *   - tensors are NHWC fp32 (Keras layout, utils/preprocessing.py:21-27);
 *   - `stream` is a cudaStream_t passed as void*; NULL = legacy default stream;
 *   - no allocation and no host synchronisation inside hot calls (workspaces are sized
 *     at s2s_unet_create from max_batch); a handle is not thread-safe, distinct handles
 *     are independent;
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with
 *     S2S_ERR_CUDA.
 */
#ifndef S2S_UNET_H
#define S2S_UNET_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S2S_ABI_VERSION 2

typedef enum {
    S2S_OK = 0,
    S2S_ERR_INVALID = -1,   /* bad argument / unsupported shape (e.g. H not divisible by 2^n_blocks) */
    S2S_ERR_CUDA = -2,      /* CUDA runtime error, see s2s_last_error() */
    S2S_ERR_NOMEM = -3,
    S2S_ERR_STATE = -4      /* call order violated (e.g. backward before forward) */
} s2s_status;

typedef enum { S2S_POOL_AVG = 0, S2S_POOL_MAX = 1 } s2s_pool_kind;
typedef enum { S2S_HEAD_SOFTMAX3 = 0, S2S_HEAD_RELU1 = 1 } s2s_head_kind;
typedef enum { S2S_LOSS_CCE = 0, S2S_LOSS_MASKED_MSE = 1 } s2s_loss_kind;
typedef enum { S2S_PREC_FP32 = 0, S2S_PREC_BF16_TC = 1, S2S_PREC_TF32 = 2 } s2s_precision;
typedef enum { S2S_ACT_ELU = 0, S2S_ACT_RELU = 1 } s2s_act_kind;   /* down()/up() `activation=` (deep_nn_models.py:139,152) */

/* Model hyper-parameters: Unet.__init__ (utils/deep_nn_models.py:19-45) + build_model's
 * dg_train_shape / output (utils/deep_nn_models.py:73-105). */
typedef struct {
    int32_t H, W, Cin;      /* dg_train_shape = (H, W, Cin) */
    int32_t filters;        /* Unet(filters=)   default 2 */
    int32_t n_blocks;       /* Unet(n_blocks=)  3..5 */
    int32_t ct_kernel;      /* Unet(ct_kernel=(k,k)) k in {2,3,5}; ct_stride is fixed (2,2) */
    int32_t pool;           /* s2s_pool_kind: apool=True -> AVG (deep_nn_models.py:148) */
    int32_t bn;             /* Unet(bn=) */
    int32_t head;           /* s2s_head_kind: output="proba" | "deterministic" */
    int32_t max_batch;      /* largest N any call will pass */
    float   bn_eps;         /* Keras BatchNormalization default 1e-3 */
    float   bn_momentum;    /* Keras BatchNormalization default 0.99 */
    int32_t precision;      /* s2s_precision: FP32 (parity path) | BF16_TC: inference forward of the thick layers
                               (Cin % 32 == 0, Cout % 16 == 0) on tcgen05 tensor cores, bf16 operands / fp32 accumulate |
                               TF32: forward AND input-gradient convolutions of every layer with Cin % 8 == 0 on tcgen05
                               (kind::tf32 on the fp32 activations, single pass, rel-L2 ~5e-4), training and inference */
    int32_t act;            /* s2s_act_kind of the hidden Conv2D layers: ELU (the reference's only setting,
                               deep_nn_models.py:139,152) | RELU (north_star "conv3x3+BatchNorm+ReLU" variant) */
} s2s_unet_cfg;

/* One named tensor of the flat parameter / state arenas (Keras kernel order). */
typedef struct {
    char    name[48];       /* Keras layer name + "/kernel" | "/bias" | "/gamma" | ... */
    int32_t arena;          /* 0 = trainable parameter arena, 1 = BN moving-statistics arena */
    int32_t ndim;
    int32_t shape[4];       /* Conv2D (kh,kw,Cin,Cout); Conv2DTranspose (kh,kw,Cout,Cin) */
    int64_t offset;         /* element offset into the arena */
    int64_t count;
} s2s_tensor_desc;

/* Adam hyper-parameters, Keras-3 form (keras optimizers.Adam; training.py:66,95). */
typedef struct {
    double lr, beta1, beta2, eps;   /* doubles so that 1-beta and the bias correction match Keras' Python floats */
} s2s_adam_cfg;

typedef struct s2s_unet s2s_unet;   /* opaque handle = one Keras `Model` (training.py:58-60,91-93) */

/* ---- library / device plumbing ------------------------------------------------------ */
int         s2s_version(void);
const char* s2s_last_error(void);
int  s2s_device_count(int* n);
int  s2s_set_device(int dev);
int  s2s_get_device(int* dev);                            /* the calling THREAD's current device */
int  s2s_stream_create(void** stream);
int  s2s_stream_destroy(void* stream);
int  s2s_stream_sync(void* stream);
/* host-layer device buffers from a per-device size-class cache (cudaFree costs milliseconds in a process holding many graphs).
 * s2s_dev_free does NOT synchronise: the caller must have synchronised the stream(s) that used the block before freeing it. */
int  s2s_dev_alloc(void** p_dev, size_t bytes);
int  s2s_dev_free(void* p_dev);
int  s2s_host_alloc(void** p_host, size_t bytes);          /* pinned */
int  s2s_host_free(void* p_host);
int  s2s_memcpy_h2d(void* dst_dev, const void* src_host, size_t bytes, void* stream);
int  s2s_memcpy_d2h(void* dst_host, const void* src_dev, size_t bytes, void* stream);
int  s2s_memcpy_d2d(void* dst_dev, const void* src_dev, size_t bytes, void* stream);
int  s2s_memset_dev(void* dst_dev, int byte, size_t bytes, void* stream);
int  s2s_event_create(void** ev);
int  s2s_event_destroy(void* ev);
int  s2s_event_record(void* ev, void* stream);
int  s2s_event_elapsed_ms(void* ev_start, void* ev_stop, float* ms);  /* syncs on ev_stop */
int  s2s_l2_flush(void* scratch_dev, size_t bytes, void* stream);      /* writes `bytes` (> L2) */

/* Per-launch profiler used by bench.py for the roofline line: while enabled every kernel is
 * bracketed by CUDA events on its launch stream (graphs are bypassed) and tagged with its
 * algorithmic bytes / flops.  s2s_prof_report writes "tag,launches,total_ms,bytes,flops" lines. */
int  s2s_prof_enable(int on);
int  s2s_prof_report(char* buf, size_t buflen);
int  s2s_prof_null_us(float* us_out, void* stream);      /* the bracket's own cost around an empty kernel */
/* measured fp32 CUDA-core peaks in TFLOP/s (the denominator of the "ffma" roofline): scalar FFMA and packed FFMA2 */
int  s2s_ffma_peak(float* scalar_tflops, float* packed_tflops, void* stream);

/* ---- model handle --------------------------------------------------------------------
 * replaces Unet(...).build_model(input_shape)      utils/training.py:58-60, 91-93
 *          (graph in utils/deep_nn_models.py:73-163) */
int  s2s_unet_create(const s2s_unet_cfg* cfg, s2s_unet** out);
int  s2s_unet_destroy(s2s_unet* h);   /* the caller drains the stream(s) it used with h first (the pool is cached) */
int  s2s_unet_param_layout(const s2s_unet* h, s2s_tensor_desc* descs, int* n);  /* descs may be NULL to query n */
int  s2s_unet_params(s2s_unet* h, float** params_dev, size_t* n);        /* trainable arena (borrowed) */
int  s2s_unet_state(s2s_unet* h, float** state_dev, size_t* n);          /* BN moving mean/var (borrowed) */
int  s2s_unet_grad_arena(s2s_unet* h, float** grads_dev, size_t* n);     /* dense grads (borrowed; NCCL all-reduce target) */
int  s2s_unet_opt_state(s2s_unet* h, float** m_dev, float** v_dev, int64_t** step_dev);
int  s2s_unet_io_buffers(s2s_unet* h, float** x_dev, float** y_dev);     /* staging [max_batch,H,W,Cin] / [max_batch,H,W,Cout] */
/* stats: {mean loss, accuracy} of the last batch; stats_acc: running {sum loss*pixels, correct, pixels}
 * since the last reset = the per-epoch 'loss' / 'accuracy' / 'val_loss' of history.history (training.py:106) */
int  s2s_unet_stats_buffers(s2s_unet* h, float** stats_dev, double** stats_acc_dev);
int  s2s_unet_reset_epoch_stats(s2s_unet* h, void* stream);
int  s2s_unet_set_lr(s2s_unet* h, double lr, void* stream);
int  s2s_unet_activation(s2s_unet* h, const char* layer_name, float** act_dev,
                         int* Hl, int* Wl, int* Cl, int* ld);            /* named Keras layer output of the last forward */
int  s2s_unet_launch_count(const s2s_unet* h, int64_t* n_kernels);       /* kernels enqueued so far (bench: gpu_launches) */
int  s2s_unet_set_graphs(s2s_unet* h, int enable);                       /* CUDA-graph replay of train/forward steps */

/* replaces model.compile(optimizer=Adam(lr), loss="categorical_crossentropy")  training.py:66-67,95-96 */
int  s2s_unet_compile(s2s_unet* h, const s2s_adam_cfg* adam, int loss_kind);

/* replaces model.predict(X) / the forward half of fit    training.py:102,133-135
 * training=0: BN uses moving statistics; training=1: batch statistics + moving update.
 * probs_dev: [N,H,W,3] (softmax head) or [N,H,W,1] (relu head). */
int  s2s_unet_forward(s2s_unet* h, const float* x_dev, int N, float* probs_dev, int training, void* stream);

/* one optimiser step of model.fit (training.py:102-103): fwd (BN batch stats) -> loss ->
 * bwd -> Keras-form Adam.  mask_dev (uint8 [H,W], nullable) is only read by MASKED_MSE.
 * stats_dev (nullable) receives {mean loss, accuracy} as 2 floats. */
int  s2s_unet_train_step(s2s_unet* h, const float* x_dev, const float* y_dev, const uint8_t* mask_dev,
                         int N, float* stats_dev, void* stream);
/* the same step from HOST batches in one call (what Keras' train_on_batch / one fit step does end to end): H2D of
 * x [N,H,W,Cin] and y [N,H,W,Cout] (pinned memory for a truly asynchronous copy), the fused step, D2H of
 * {mean loss, accuracy} into stats_host (nullable) and a stream synchronisation. */
int  s2s_unet_train_step_host(s2s_unet* h, const float* x_host, const float* y_host, int N, float* stats_host, void* stream);
/* a stream of nsteps such steps, batch i from x_hosts[i] / y_hosts[i], {loss, accuracy} of step i into
 * stats_host[2 i .. 2 i + 1] (nullable): what model.fit does with host arrays.  Every step still copies its own batch H2D and
 * its own result D2H; the copy of batch i + 1 is staged on a second stream while step i computes.  One synchronisation per 256
 * steps.  Pinned (page-locked / registered) batches are copied directly; ordinary pageable memory — the reference's NumPy arrays —
 * is first copied into a ring of pinned slots by the calling thread while the GPU computes (a pageable source would turn the
 * asynchronous copy into a blocking one).  The host arrays may be released as soon as the call returns. */
int  s2s_unet_train_steps_host(s2s_unet* h, const float* const* x_hosts, const float* const* y_hosts, int nsteps, int N,
                               float* stats_host, void* stream);
/* fwd + loss + bwd only: leaves dense grads in the grad arena (for NCCL all-reduce by the
 * host); grad_scale multiplies the loss gradient (1/world_size under data parallel). */
int  s2s_unet_backward_only(s2s_unet* h, const float* x_dev, const float* y_dev, const uint8_t* mask_dev,
                            int N, float grad_scale, float* stats_dev, void* stream);
/* Adam on the handle's arenas using the (all-reduced) dense grad arena. */
int  s2s_unet_apply_adam(s2s_unet* h, void* stream);
/* loss / accuracy of a batch in inference mode (the validation pass of fit, training.py:102). */
int  s2s_unet_eval_batch(s2s_unet* h, const float* x_dev, const float* y_dev, const uint8_t* mask_dev,
                         int N, float* stats_dev, void* stream);
/* Grad-CAM (Selvaraju et al.) for class `cls` at the named Keras layer: out [N,Hl,Wl]. */
int  s2s_unet_gradcam(s2s_unet* h, const float* x_dev, int N, const char* layer_name, int cls,
                      float* cam_dev, void* stream);

/* On-device fit protocol: the data set stays resident in HBM; one call
 * enqueues a whole epoch of model.fit (training.py:102-103) / the validation pass / model.predict
 * (training.py:133-135) without host round trips.  perm_dev [T] (nullable = identity) is the
 * epoch's shuffled sample order; the partial last batch is kept.  Loss/accuracy accumulate in
 * the stats_acc buffer (s2s_unet_stats_buffers). */
int  s2s_unet_fit_epoch(s2s_unet* h, const float* x_all_dev, const float* y_all_dev, const int32_t* perm_dev,
                        int T, int batch_size, const uint8_t* mask_dev, void* stream);
int  s2s_unet_eval_dataset(s2s_unet* h, const float* x_all_dev, const float* y_all_dev, int T, int batch_size,
                           const uint8_t* mask_dev, void* stream);
int  s2s_unet_predict_dataset(s2s_unet* h, const float* x_all_dev, int T, int batch_size, float* probs_all_dev,
                              void* stream);

/* ---- batch-sharded data parallelism over NVLink peer memory ----------------------
 * The reference trains on one device (model.fit, training.py:102).  Splitting its batch over the GPUs of
 * one node needs a gradient exchange per optimiser step and, for parity with the single-device batch, a
 * BatchNormalization statistics exchange per layer and direction.  Both are done by this library's own kernels
 * with loads/stores on peer memory (CUDA IPC), fused with the Adam update / the BN finalisation; the host
 * (one process per GPU) only exchanges the 64-byte IPC handles once, e.g. with torch.distributed.
 *   s2s_dp_create      allocates this rank's exchange buffer for a flat arena of n_floats parameters
 *   s2s_dp_ipc_handle  writes its cudaIpcMemHandle_t (64 bytes) to handle64 (host memory)
 *   s2s_dp_connect     maps the peers' buffers; handles = world x 64 bytes in rank order (host memory)
 *   s2s_dp_error       0, or 1 + sync group after a peer timed out (synchronous device read)
 *   s2s_unet_attach_dp binds a model replica to the communicator; sync_bn != 0 = global batch statistics
 *   s2s_unet_dp_train_step  fwd -> loss -> bwd on this rank's n_local samples of a global batch of n_global,
 *                      then ONE kernel: all-reduce over peer memory (fixed rank order) + Keras-form Adam.
 *                      stats_dev (nullable) receives the global {mean loss, accuracy}. */
typedef struct s2s_dp s2s_dp;
int  s2s_dp_create(int rank, int world, size_t n_floats, s2s_dp** out);
int  s2s_dp_ipc_handle(s2s_dp* d, void* handle64);
int  s2s_dp_connect(s2s_dp* d, const void* handles);
int  s2s_dp_error(s2s_dp* d, int* err);
int  s2s_dp_destroy(s2s_dp* d);
int  s2s_unet_attach_dp(s2s_unet* h, s2s_dp* d, int sync_bn);   /* d may be NULL to detach */
int  s2s_unet_dp_train_step(s2s_unet* h, const float* x_dev, const float* y_dev, int n_local, int n_global,
                            float* stats_dev, void* stream);
/* the same from HOST shards in one call (H2D, step, D2H of the global {loss, accuracy}, synchronise), like
 * s2s_unet_train_step_host */
int  s2s_unet_dp_train_step_host(s2s_unet* h, const float* x_host, const float* y_host, int n_local, int n_global,
                                 float* stats_host, void* stream);

/* a stream of nsteps data-parallel steps from host shards (like s2s_unet_train_steps_host; stats = the GLOBAL {loss, accuracy}) */
int  s2s_unet_dp_train_steps_host(s2s_unet* h, const float* const* x_hosts, const float* const* y_hosts, int nsteps, int n_local,
                                  int n_global, float* stats_host, void* stream);

/* ---- stand-alone fused Adam (Keras-3 form) on caller arenas ------------------------- */
int  s2s_adam_step(float* p_dev, const float* g_dev, float* m_dev, float* v_dev, size_t n,
                   const s2s_adam_cfg* cfg, int64_t step /* 1-based */, void* stream);

/* ---- skill reductions ------------------------------------------------------------------
 * RPS per gridpoint: replaces xskillscore.rps(obs, fcst, dim='T', input_distributions='p')
 * (utils/performance_metrics.py:26-40).  p,o: [T,Y,X,3]; NaN in o[...,0] marks a missing
 * observation (skipped).  out: [Y,X]. */
int  s2s_rps_map(const float* p_dev, const float* o_dev, int T, int Y, int X, float* out_dev, void* stream);
/* RPSS = 1 - RPS_f / RPS_ref (utils/performance_metrics.py:44-45). */
int  s2s_rpss_map(const float* fcst_dev, const float* ref_dev, const float* o_dev, int T, int Y, int X,
                  float* out_dev, void* stream);
/* CC = xr.corr(x,y,'T'); ACC = xr.corr of ISO-week anomalies (ACCs.ipynb:362-388).
 * order [T]: start indices sorted by ISO-week group; group_start [n_groups+1]: offsets of each
 * group in order[].  x,y: [T,Y,X]; NaN pairs are skipped; acc/cc: [Y,X] (either may be NULL). */
int  s2s_acc_map(const float* x_dev, const float* y_dev, const int32_t* order_dev, const int32_t* group_start_dev,
                 int n_groups, int T, int Y, int X, float* acc_dev, float* cc_dev, void* stream);
/* ensemble-mean predictor image: x [T,M,Y,X] -> [T,Y,X]  (utils/preprocessing.py:21-23) */
int  s2s_ensemble_mean(const float* x_dev, int T, int M, int Y, int X, float* out_dev, void* stream);
/* MME combine: mean over models of probs then renormalise over category (training.py:344-350).
 * probs: [n_models][T,Y,X,3] contiguous. */
int  s2s_mme_combine(const float* probs_dev, int n_models, int64_t n_points, float* out_dev, void* stream);

/* ---- predictand pre-processing -------------------------------------------------
 * ISO-week rolling tercile edges: replaces the per-week `observations_weekly.quantile([1/3, 2/3], dim='T')` loop of
 * rolling_labeler (utils/preprocessing.py:112-126; also make_tercile_labeler :10 with one window).  y: [T, YX]
 * (float32, or float64 when is_f64) training predictand; window w covers the starts win_idx[win_start[w] ..
 * win_start[w+1]) (int32, built by the host from the ISO weeks); NaNs are skipped (nanquantile), arithmetic follows
 * numpy's 'linear' method bit for bit; edges: [n_weeks][2][YX] float64 (NaN for an all-NaN point). */
int  s2s_tercile_edges(const void* y_dev, int is_f64, const int32_t* win_start_dev, const int32_t* win_idx_dev,
                       int n_weeks, int64_t YX, int max_window_len, double* edges_dev, void* stream);
/* Labels 0 (y < e0) / 2 (y > e1) / 1, NaN where an edge is NaN (utils/preprocessing.py:137-158), and/or their
 * to_categorical(., 3) one-hot (utils/preprocessing.py:426-428).  week_slot[t] = row of edges to use for start t
 * (nearest training week).  labels: [T, YX] float32, onehot: [T, YX, 3] float32; either may be NULL. */
int  s2s_tercile_label(const void* y_dev, int is_f64, const int32_t* week_slot_dev, const double* edges_dev,
                       int T, int64_t YX, float* labels_dev, float* onehot_dev, void* stream);

/* ---- extended logistic regression baseline ----------------------------------------
 * Replaces the Python Y x X loop around statsmodels.GLM(Binomial).fit() of train_single_bootstrap_ELR
 * (utils/training.py:402-530): one 3-parameter IRLS fit per gridpoint on the 2T rows (start, threshold in {33, 67})
 * with design [1, ensemble-mean forecast, threshold], response [y <= tercile edge] (rolling_labeler_ELR,
 * utils/preprocessing.py:270-333), then P(below/normal/above) for the training and the test starts.
 * x_*: [T, YX] float64; y_train: [T, YX] float32 | float64; slot_*[t]: row of edges for start t; edges:
 * [n_weeks][2][YX] from s2s_tercile_edges; p_*: [T, YX, 3] float64 (NaN for skipped gridpoints, 1/3 for the dropped
 * starts of a fitted gridpoint); iters (nullable): IRLS updates per gridpoint. */
int  s2s_elr_fit_predict(const double* x_train_dev, const void* y_train_dev, int y_is_f64, const int32_t* slot_train_dev,
                         const double* x_test_dev, const int32_t* slot_test_dev, const double* edges_dev, int T, int Tt,
                         int64_t YX, double* p_train_dev, double* p_test_dev, int32_t* iters_dev, void* stream);

/* ---- single-operator entry points (used by the parity tests and by profiling) --------- */
/* y = ELU(conv3x3_same(x, w) + b)   Conv2D(3x3, elu, same)  deep_nn_models.py:142,145,157,160 */
int  s2s_op_conv3x3_fwd(const float* x_dev, const float* w_dev, const float* b_dev, float* y_dev,
                        int N, int H, int W, int Cin, int Cout, int apply_elu, void* stream);
/* same operator on the tensor cores: bf16 operands (cast inside), fp32 accumulation in TMEM via tcgen05.mma, tiles
 * staged by TMA; needs Cin % 64 == 0, Cout % 16 == 0, Cout <= 256.  bf16 tolerance (rel-L2 <= 1e-2). */
int  s2s_op_conv3x3_fwd_tc(const float* x_dev, const float* w_dev, const float* b_dev, float* y_dev,
                           int N, int H, int W, int Cin, int Cout, int apply_elu, void* stream);
/* the same operator and its input gradient on the tcgen05 tensor cores with tf32 operands read straight from the fp32
 * tensors (csrc/tc3conv.cuh): npass = 1 single pass (rel-L2 ~5e-4), npass = 3 error-compensated 3xTF32 split (rel-L2
 * ~1e-6, the fp32 parity bar).  Needs Cin % 8 == 0 (forward) / Cout % 8 == 0 (dgrad) and the other count % 4 == 0. */
int  s2s_op_conv3x3_fwd_tf32(const float* x_dev, const float* w_dev, const float* b_dev, float* y_dev,
                             int N, int H, int W, int Cin, int Cout, int apply_elu, int npass, void* stream);
int  s2s_op_conv3x3_dgrad_tf32(const float* dz_dev, const float* w_dev, const float* act_dev, float* dx_dev,
                               int N, int H, int W, int Cin, int Cout, int npass, void* stream);
/* dx = conv3x3_dgrad(dz, w) [* ELU'(act)]  (act nullable) */
int  s2s_op_conv3x3_dgrad(const float* dz_dev, const float* w_dev, const float* act_dev, float* dx_dev,
                          int N, int H, int W, int Cin, int Cout, void* stream);
/* dw [3,3,Cin,Cout], db [Cout] */
int  s2s_op_conv3x3_wgrad(const float* x_dev, const float* dz_dev, float* dw_dev, float* db_dev,
                          int N, int H, int W, int Cin, int Cout, void* stream);
/* the same gradients as a pixel-contraction GEMM on the tcgen05 tensor cores (csrc/tcwgrad.cuh; precision = tf32, rel-L2
 * ~7e-4): needs Cin % 4 == 0 and Cout % 4 == 0.  x / dz hold n_max >= N images, only the first N contribute. */
int  s2s_op_conv3x3_wgrad_tf32(const float* x_dev, const float* dz_dev, float* dw_dev, float* db_dev,
                               int N, int H, int W, int Cin, int Cout, int n_max, void* stream);
/* Conv2DTranspose(k, strides 2, same): x [N,h,w,Cin] -> y [N,2h,2w,Cout]; w (k,k,Cout,Cin)  deep_nn_models.py:154 */
int  s2s_op_convt_fwd(const float* x_dev, const float* w_dev, const float* b_dev, float* y_dev,
                      int N, int h, int w, int Cin, int Cout, int k, void* stream);
int  s2s_op_convt_dgrad(const float* dy_dev, const float* w_dev, float* dx_dev,
                        int N, int h, int w, int Cin, int Cout, int k, void* stream);
/* the transposed conv and its input gradient on the tcgen05 tensor cores (csrc/tc3conv.cuh, single tf32 pass, rel-L2
 * ~5e-4): forward = four stride-1 3x3 convs on the input grid, one per output parity; dgrad = one conv contracting over the
 * four stride-2 parity planes of dy.  Needs Cin % 8 == 0 and Cout % 8 == 0. */
int  s2s_op_convt_fwd_tf32(const float* x_dev, const float* w_dev, const float* b_dev, float* y_dev,
                           int N, int h, int w, int Cin, int Cout, int k, void* stream);
int  s2s_op_convt_dgrad_tf32(const float* dy_dev, const float* w_dev, float* dx_dev,
                             int N, int h, int w, int Cin, int Cout, int k, void* stream);
/* dw (k,k,Cout,Cin) of the transposed conv as a pixel-contraction GEMM over the four parity planes of dy (csrc/tcwgrad.cuh);
 * x / dy hold n_max >= N images, only the first N contribute.  Needs Cin % 4 == 0 and Cout % 4 == 0. */
int  s2s_op_convt_wgrad_tf32(const float* x_dev, const float* dy_dev, float* dw_dev,
                             int N, int h, int w, int Cin, int Cout, int k, int n_max, void* stream);
int  s2s_op_convt_wgrad(const float* x_dev, const float* dy_dev, float* dw_dev, float* db_dev,
                        int N, int h, int w, int Cin, int Cout, int k, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* S2S_UNET_H */
