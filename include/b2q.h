/* b2q.h -- C ABI of libb2q.so: B200 (sm_100a) kernels for the int8 fake-quantization CustomOps of
 * XiaotaoChen/resnet.mxnet.
 *
 * This is the drop-in boundary.  Every entry point takes plain device pointers, sizes and a CUDA stream
 * (as void*); no framework types cross it.  The library never allocates or frees tensor memory and never
 * synchronises the device: all work is enqueued on the caller's stream.  Tensors are float32, dense,
 * row-major (NCHW activations, (Cout, Cin/g, kh, kw) weights) exactly as the reference's NDArrays.
 *
 * Each function names the reference code it replaces (paths relative to the reference repo root).
 * Return value: 0 on success, non-zero on failure with a message in b2q_last_error() (thread-local).
 *
 * "Segmented" tensors: several entry points view a tensor as (outer, groups, inner), element (o, g, i) at
 * ((o * groups) + g) * inner + i, with one threshold per g:
 *     whole tensor                     outer=1      groups=1          inner=N
 *     per-out-channel weight           outer=1      groups=Cout       inner=Cin/g*kh*kw
 *     GDRQ grouped weight              outer=1      groups=Cout/gs    inner=gs*Cin/g*kh*kw
 *     GDRQ grouped activation (NCHW)   outer=N      groups=C/gs       inner=gs*H*W
 * which replaces the reference's swapaxes/reshape/broadcast_like chains (core/operator/GDRQ.py:88-118,
 * symbol/quant_ops.py:20-24) without moving data.
 */
#ifndef B2Q_H_
#define B2Q_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2Q_ABI_VERSION 1

/* MXNet OpReqType (include/mxnet/op_attr_types.h [upstream]); CustomOp.assign() semantics. */
#define B2Q_REQ_NULL 0
#define B2Q_REQ_WRITE 1
#define B2Q_REQ_INPLACE 2
#define B2Q_REQ_ADD 3

typedef struct b2q_ctx b2q_ctx; /* per-device handle: SM count + a small ring of reduction workspaces */

/* ---- lifecycle / errors ------------------------------------------------------------------------ */
int b2q_abi_version(void);
const char* b2q_last_error(void);
int b2q_create(int device, b2q_ctx** out);
int b2q_destroy(b2q_ctx* ctx);
int b2q_num_sms(b2q_ctx* ctx);
/* cudaStreamSynchronize(stream) on ctx's device.  The library never synchronises by itself; this is for hosts whose
 * scheduler cannot see the library's launches (MXNet's engine: a CustomOp callback must not return before its kernels
 * are done, python/mxnet/operator.py [upstream]) -- the MXNet-side operator wrapper calls it with the legacy default
 * stream (0) after forward / backward, see INTEGRATION.md section 3.                                                */
int b2q_stream_synchronize(b2q_ctx* ctx, void* stream);
/* run-time knobs for benchmarking sweeps: "blocks_per_sm" (grid = SMs x this), "reverse" (QDQ sweep walks
 * descending addresses to reuse what the reduction left in L2), "fast_div" (reciprocal fast path on/off),
 * "peer_reduce_blocks_per_sm" (grid of the max reduction in the peer-memory exchange), "pdl" (programmatic dependent
 * launch between consecutive whole-tensor kernels on/off), "timing" (see b2q_timing_read), "resident" (single-launch
 * forward for tensors that fit on chip), "peer_mode" (1: ticket-free exchange kernels, 0: the first-generation ones,
 * 2 / 3: further tiles staged in shared memory during the wait, 4: the reduction's last block publishes),
 * "peer_stage_early", "peer_publish_blocks_per_sm", "peer_allreduce_blocks" / "peer_allreduce_blocks_per_sm" (grid of
 * b2q_peer_allreduce_*), "peer_timeout_ms", "host_ste_copy" (host-buffer straight-through backward copied host to host
 * instead of through the GPU), "dorefa_tanh_max" (DoReFa: element-wise max of |tanh| instead of tanhf(max|w|)),
 * "cluster_fwd" / "cluster_max_elems" / "cluster_max_elems_mean" / "cluster_words_per_cta" (single-launch cluster
 * forward for small tensors), "reverse_min_mb", "bn_variant" / "bn_pieces_per_sm" (batch-statistics kernel),
 * "stream_reduce" / "stream_stages" (TMA-staged reduction ring), "stream_icvt" (float -> double conversions on the
 * integer pipe).  Results never depend on them (the mean-based statistics to within the order of an exact double sum). */
int b2q_set_option(b2q_ctx* ctx, const char* key, int value);
int b2q_get_option(b2q_ctx* ctx, const char* key, int* value);
/* number of kernels this library has launched through ctx since creation (bench.py's gpu_launches) */
int64_t b2q_launch_count(b2q_ctx* ctx);
/* With option "timing"=1 every launch of the whole-tensor kernels is bracketed by CUDA events on the caller's
 * stream.  b2q_timing_read sums them per kind (1 reduction, 2 QDQ sweep, 3 STE backward, 4 masked backward,
 * 5 segmented/other, 0 all): device milliseconds, ALGORITHMIC bytes (4, 8, 8, 12 B/element) and launch count;
 * reset!=0 clears the records.  Synchronises on the recorded events.  Not usable under stream capture.      */
int b2q_timing_read(b2q_ctx* ctx, int kind, double* total_ms, double* total_bytes, int64_t* count, int reset);
/* the same, restricted to launches whose algorithmic bytes lie in [min_bytes, max_bytes) (max_bytes <= 0: no upper
 * bound): lets a benchmark quote large tensors apart from the launch-bound small ones                          */
int b2q_timing_read_range(b2q_ctx* ctx, int kind, double min_bytes, double max_bytes, double* total_ms,
                          double* total_bytes, int64_t* count, int reset);

/* ---- primitives ---------------------------------------------------------------------------------
 * K1/K2  b2q_absmax_f32    stat[g] = max |x|          replaces mx.nd.abs -> mx.nd.max
 *        (symbol/quant_ops.py:18-26,34-35; symbol/clip_grad_quantization_int8.py:23-34,39-40)
 * K7     b2q_meanabs_f32   stat[g] = fl(sum|x| / n_g)  replaces mx.nd.abs -> mx.nd.mean
 *        (core/operator/GDRQ.py:67-72,95-100; symbol/fold_bn_v1_gdrq.py:56-58,78-90)
 * Both are deterministic (fixed partition, fixed combination order).                                */
int b2q_absmax_f32(b2q_ctx* ctx, const float* x, int64_t outer, int64_t groups, int64_t inner,
                   float* stat, void* stream);
int b2q_meanabs_f32(b2q_ctx* ctx, const float* x, int64_t outer, int64_t groups, int64_t inner,
                    float* stat, void* stream);

/* K3  threshold update from an already reduced statistic (used after a cross-rank allreduce(max)).
 * mode: B2Q_UPD_*.  aux is updated in place; p0/p1 are (ema_decay, 1-ema_decay) or (ktimes, lamda).
 * clip_out (may be NULL) receives the batch threshold where the op clips with it
 * (symbol/fold_bn_v1_gdrq.py:58,67).                                                                */
#define B2Q_UPD_STORE 1        /* aux = stat                         quant_ops.py:31, clip_grad...py:26,34 */
#define B2Q_UPD_EMA 2          /* aux = aux*d + stat*(1-d)           quant_ops.py:37, clip_grad...py:46    */
#define B2Q_UPD_GDRQ_WEIGHT 3  /* aux = k*stat                       GDRQ.py:72-74                          */
#define B2Q_UPD_GDRQ_ACT 4     /* aux = aux + lamda*(aux - k*stat)   GDRQ.py:72,76                          */
#define B2Q_UPD_TWICE_STORE 5  /* aux = 2*stat                       fold_bn_v1_gdrq.py:58,61 / :82,90,95   */
#define B2Q_UPD_TWICE_EMA 6    /* aux = aux*d + (2*stat)*(1-d)       fold_bn_v1_gdrq.py:58,64               */
int b2q_threshold_update_f32(b2q_ctx* ctx, int mode, const float* stat, float* aux, float* clip_out,
                             int64_t groups, float p0, float p1, void* stream);

/* K4/K8  one-sweep clip + quantize-dequantize:  y = fl(roundf(fl(c(x) / q)) * q),  q = fl(thr[g]/qlevel)
 * (qlevel <= 0: thr holds q itself).  Bit-exact with the reference's x/q -> round -> *q chain
 * (quant_ops.py:28,40; clip_grad...py:36,48-51; GDRQ.py:79-86,109-114).
 * clip_mode: B2Q_CLIP_*; clip thresholds come from clip_thr (NULL: same as thr).
 * do_round=0 gives the clip-only output of GDRQ's delay_quant branch (GDRQ.py:81-82).
 * codes (may be NULL): int32 side output of roundf(c(x)/q) for parity checks.
 * prescale (may be NULL) multiplies row (o*groups+g) by gamma/sqrt(var+eps) before anything else
 * (fold_bn_v1_gdrq.py:72-74): pass gamma, var (length outer*groups) and eps.                        */
#define B2Q_CLIP_NONE 0
#define B2Q_CLIP_SYM 1        /* mx.nd.clip(x, -T, T)                          */
#define B2Q_CLIP_WHERE_LE 2   /* where(|x| <= T, x, T*sign(x))   GDRQ.py:109   */
#define B2Q_CLIP_ZERO_T 3     /* mx.nd.clip(x, 0, T)             GDRQ.py:202   */
#define B2Q_CLIP_PACT 4       /* where(x < T, x, T)              PACT.py:125   */
#define B2Q_CLIP_WHERE_LT 5   /* where(|x| < T, x, T*sign(x))    PACT.py:193   */
int b2q_qdq_f32(b2q_ctx* ctx, const float* x, float* y, int64_t outer, int64_t groups, int64_t inner,
                const float* thr, const float* clip_thr, float qlevel, int clip_mode, int do_round,
                int req, int32_t* codes, const float* prescale_gamma, const float* prescale_var,
                float prescale_eps, void* stream);

/* True-int8 export (inference): codes[i] = clamp(roundf(c(x[i]) / q), -128, 127) as int8 and steps[g] = q[g] =
 * fl(thr[g] / qlevel), i.e. exactly the integers the QDQ sweep multiplies by q, so a
 * convolution can consume them on int8 / fp8 tensor cores.  steps may be NULL.                               */
int b2q_export_int8_f32(b2q_ctx* ctx, const float* x, int8_t* codes, float* steps, int64_t outer, int64_t groups,
                        int64_t inner, const float* thr, float qlevel, int clip_mode, void* stream);

/* K5  straight-through backward: in_grad (req) out_grad      quant_ops.py:41-42, GDRQ.py:126        */
int b2q_ste_bwd_f32(b2q_ctx* ctx, const float* dy, float* dx, int64_t n, int req, void* stream);

/* in_grad[i][:] = 0 for the inputs that get no gradient (symbol/fold_bn_v1_gdrq.py:124-125): a memset on the caller's
 * stream.                                                                                              */
int b2q_zero_f32(b2q_ctx* ctx, float* x, int64_t n, void* stream);

/* K6  masked backward, one pass:  dx (req) dy * mask(x, thr[g])
 * mask_mode B2Q_MASK_OPEN   [x > -T][x < T]   clip_grad_quantization_int8.py:61-67
 *           B2Q_MASK_ABS_LE [|x| <= T]        GDRQ.py:132-133,145-148
 *           B2Q_MASK_LT     [x < T]           GDRQ.py:207-208
 * thr == NULL: thr_imm is used (CLIP_RELU_PY's constant threshold).                                  */
#define B2Q_MASK_OPEN 1
#define B2Q_MASK_ABS_LE 2
#define B2Q_MASK_LT 3
int b2q_mask_bwd_f32(b2q_ctx* ctx, const float* x, const float* dy, float* dx, int64_t outer,
                     int64_t groups, int64_t inner, const float* thr, float thr_imm, int mask_mode,
                     int req, void* stream);

/* ---- fused per-operator entry points (what the Python CustomOp bodies call) ---------------------- */

/* Quantization_int8.forward, non-delay branch          symbol/quant_ops.py:17-40
 * ClipGrad_Quantization_int8.forward, non-delay branch symbol/clip_grad_quantization_int8.py:19-51
 * variant 0 = Quantization_int8_V2, 1 = ClipGrad_Quantization_int8.
 * aux: minmax state [1] or [rows] (per-channel weight).  rows*cols = numel.  init: ClipGrad first-batch flag.
 * Whole-tensor case = two launches: a reduction that publishes max|x| with one tagged atomicMax per block, and a
 * QDQ sweep whose blocks derive the (EMA-updated) threshold in registers and whose block 0 writes aux.
 * For data-parallel training b2q_minmax_quant_stat_f32 only reduces (max|x| -> stat_out[groups]) so the caller can
 * allreduce(max) across ranks and finish with b2q_minmax_quant_finish_f32 (or see b2q_peer_* below).    */
int b2q_minmax_quant_fwd_f32(b2q_ctx* ctx, int variant, const float* x, float* y, float* aux,
                             int64_t rows, int64_t cols, int is_weight, int per_channel, int is_train,
                             int init, float ema_decay, float one_minus_decay, int req, void* stream);
int b2q_minmax_quant_stat_f32(b2q_ctx* ctx, const float* x, int64_t rows, int64_t cols, int per_channel,
                              float* stat_out, void* stream);
int b2q_minmax_quant_finish_f32(b2q_ctx* ctx, int variant, const float* x, float* y, float* aux,
                                const float* stat, int64_t rows, int64_t cols, int is_weight,
                                int per_channel, int is_train, int init, float ema_decay,
                                float one_minus_decay, int req, void* stream);
/* ClipGrad_Quantization_int8.backward (act)            symbol/clip_grad_quantization_int8.py:61-67 */
int b2q_clipgrad_bwd_f32(b2q_ctx* ctx, const float* x, const float* dy, float* dx, const float* aux,
                         int64_t n, void* stream);

/* GDRQ_PY.forward / backward                           core/operator/GDRQ.py:64-118 / :124-152
 * alpha: [groups].  (outer, groups, inner) as in the header comment; group_size==-1 -> groups=1.      */
int b2q_gdrq_fwd_f32(b2q_ctx* ctx, const float* x, float* y, float* alpha, int64_t outer, int64_t groups,
                     int64_t inner, int is_weight, int fix_alpha, int do_round, float qlevel, float ktimes,
                     float lamda, int req, void* stream);
int b2q_gdrq_bwd_f32(b2q_ctx* ctx, const float* x, const float* dy, float* dx, const float* alpha,
                     int64_t outer, int64_t groups, int64_t inner, int req, void* stream);

/* GDRQ_Fold_BN.forward up to the convolution           symbol/fold_bn_v1_gdrq.py:53-96,113
 * data_q  = QDQ(clip(data, +-2 mean|data|)) with the EMA scale in aux_data[1]      (:53-68)
 * weight_q = QDQ(clip(w * gamma/sqrt(var+eps), +-T)), T = 2 mean|w'| per tensor or per row; aux_weight=T (:70-96)
 * bias[c] = beta - mean*gamma/sqrt(var+eps)                                          (:113)
 * The convolution itself (:99-110) stays a library call (cuDNN) made by the caller.                   */
int b2q_foldbn_data_fwd_f32(b2q_ctx* ctx, const float* x, float* y, float* aux_data, int64_t n, int init,
                            float ema_decay, float one_minus_decay, void* stream);
int b2q_foldbn_weight_fwd_f32(b2q_ctx* ctx, const float* w, float* w_q, float* bias, float* aux_weight,
                              const float* gamma, const float* beta, const float* mean, const float* var,
                              float eps, int64_t cout, int64_t cols, int per_channel, int quantize,
                              int is_train, void* stream);

/* The step either side of the fold-BN operator.  In the reference's graph
 * (symbol/fold_bn_v1_gdrq.py:268-287) BatchNorm_v1(output_mean_var=True) reduces the convolution output (n, c, hw) to
 * its batch mean / variance [upstream batch_norm_v1-inl.h: mean = fl(scale * sum x), var = fl(scale * sum (x-mean)^2),
 * scale = fl(c / size)], which GDRQ_Fold_BN folds into the weight.  b2q_bn_batch_stats_f32 is that reduction;
 * b2q_bnstat_foldbn_weight_fwd_f32 does the reduction AND the weight path of b2q_foldbn_weight_fwd_f32 in one launch
 * when the weight is quantised per out-channel (the block that completes a channel folds and quantises its row),
 * in two launches otherwise.  mean / var [c] are outputs.                                                   */
int b2q_bn_batch_stats_f32(b2q_ctx* ctx, const float* y, int64_t n, int64_t c, int64_t hw, float* mean, float* var,
                           void* stream);
int b2q_bnstat_foldbn_weight_fwd_f32(b2q_ctx* ctx, const float* conv_out, int64_t n, int64_t c, int64_t hw,
                                     float* mean, float* var, const float* w, float* w_q, float* bias,
                                     float* aux_weight, const float* gamma, const float* beta, float eps,
                                     int64_t cols, int per_channel, int quantize, int is_train, void* stream);

/* QUANT_STE_PY (PACT.py:245-252) and PACT forward (PACT.py:125-128,193-198) are b2q_absmax_f32 + b2q_qdq_f32.
 * CLIP_RELU_PY.forward  core/operator/GDRQ.py:200-204: clip(x, 0, threshold) then QDQ with q (= threshold/L
 * computed by the caller in double like the reference); backward is b2q_mask_bwd_f32(B2Q_MASK_LT, thr_imm). */
int b2q_clip_relu_fwd_f32(b2q_ctx* ctx, const float* x, float* y, int64_t n, float threshold, float q,
                          int req, void* stream);
/* the remaining second-tier operators:                                                                 */
/* WNQ_PY  core/operator/WNQ.py:51-85 */
int b2q_wnq_fwd_f32(b2q_ctx* ctx, const float* x, float* y, int64_t rows, int64_t cols, int per_channel,
                    float qlevel, int req, void* stream);
int b2q_wnq_bwd_f32(b2q_ctx* ctx, const float* x, const float* dy, float* dx, int64_t rows, int64_t cols,
                    int per_channel, int req, void* stream);
/* PACT_PY / PACT_V2_PY backward  core/operator/PACT.py:142-144 / :201-203 (two_sided=1 for V2).
 * dgamma (req_gamma) sum of the gradient routed to the clipped branch.                                 */
int b2q_pact_bwd_f32(b2q_ctx* ctx, const float* x, const float* dy, float* dx, float* dgamma,
                     const float* gamma, int64_t n, int two_sided, int req, int req_gamma, void* stream);
/* DoReFa_PY  core/operator/PACT.py:44-50 / :76-77 */
int b2q_dorefa_fwd_f32(b2q_ctx* ctx, const float* x, float* y, float* vmax_out, int64_t n, float qlevel,
                       int req, void* stream);
int b2q_dorefa_bwd_f32(b2q_ctx* ctx, const float* x, const float* dy, float* dx, const float* vmax,
                       int64_t n, int req, void* stream);
/* QIL_PY / QIL_V2_PY / QIL_V3_PY  core/operator/QIL.py:50-124, QIL_V2.py:34-71, QIL_V3.py:36-70
 * variant 1/2/3; p0,p1 are the two learnable scalars in the reference's argument order
 * (pruning_point,clipping_point | center,distance | ep,ed).  V1 clamps p0>=0, p1<=1 in place (QIL.py:51-54). */
int b2q_qil_fwd_f32(b2q_ctx* ctx, int variant, const float* x, float* y, float* p0, float* p1, int64_t n,
                    float qlevel, int req, void* stream);
int b2q_qil_bwd_f32(b2q_ctx* ctx, int variant, const float* x, const float* dy, float* dx, const float* p0,
                    const float* p1, float* dp0, float* dp1, int64_t n, int req, int req_p0, int req_p1,
                    void* stream);

/* Exhaustive self tests of the two numerical shortcuts in the second-tier kernels; *failures = number of inputs on which
 * the shortcut differs from the reference arithmetic (must be 0).  which = 1: code / L by reciprocal + two FMAs versus
 * IEEE division, every integer |code| <= 4 L, L = 2^nbits - 1, nbits = 1..16.  which = 2: the facts behind DoReFa_PY's
 * max|tanh(w)| = max(tanhf(max|w|), max of tanhf over the elements in a 129-ulp window around |w| = 0.6): tanhf is odd,
 * monotonic non-decreasing over all finite positive floats outside that window, and the window's values lie between
 * tanhf of its lowest float and tanhf of the first float above it.  which = 3 / 4 (diagnostic): bit pattern of the
 * largest x with tanhf(next(x)) < tanhf(x) / with tanhf(-x) != -tanhf(x), 0 if none.  which = 5: the float32 ->
 * float64 conversion on the integer pipe that every double-precision sum uses (mean|x| of GDRQ_PY / GDRQ_Fold_BN,
 * BatchNorm_v1 statistics) against the conversion instruction over all 2^32 bit patterns.  Synchronous.          */
int b2q_selftest(b2q_ctx* ctx, int which, int64_t* failures);

/* ---- multi-tensor: every weight of a network in two launches (forward) / one launch (backward) ---------
 * The weight tensors of a network are ~1% of a step's bytes but, launched one by one (3 launches each), ~7% of its
 * time.  A plan is built once from the (stable) parameter pointers; b2q_multi_weight_quant_fwd_f32 then runs the
 * weight path of Quantization_int8_V2 (variant 0, symbol/quant_ops.py:17-31) or ClipGrad_Quantization_int8
 * (variant 1, symbol/clip_grad_quantization_int8.py:19-36) for all of them: one max|w| launch + one QDQ launch;
 * b2q_multi_weight_ste_bwd_f32 is their straight-through backward dx = dy (quant_ops.py:41-42).  Results are
 * bit-identical to the per-tensor entry points.                                                               */
typedef struct {
    const float* x;      /* weight                                   */
    float* y;            /* quantised weight                         */
    float* aux;          /* minmax state: [1] or [rows]              */
    const float* dy;     /* gradient w.r.t. y (may be NULL: no backward) */
    float* dx;           /* gradient w.r.t. x (may be NULL)          */
    int64_t rows;        /* out channels                             */
    int64_t cols;        /* elements per out channel                 */
    int32_t per_channel; /* is_weight_perchannel                     */
    int32_t reserved;
} b2q_weight_desc;
typedef struct b2q_multi_plan b2q_multi_plan;
int b2q_multi_plan_create(b2q_ctx* ctx, const b2q_weight_desc* descs, int count, b2q_multi_plan** out);
int b2q_multi_plan_destroy(b2q_ctx* ctx, b2q_multi_plan* plan);
int b2q_multi_weight_quant_fwd_f32(b2q_ctx* ctx, b2q_multi_plan* plan, int variant, int is_train, void* stream);
/* GDRQ_PY weight nodes (core/operator/GDRQ.py:69-74,97-102,109-114): aux = alpha; descriptors use per_channel=0 for
 * group_size == -1 and (rows = channels/group_size, cols = group_size * elements per channel, per_channel=1) for
 * grouped weights.  One |w| sum launch (skipped when fix_alpha) + one clip/round launch for all tensors.            */
int b2q_multi_gdrq_weight_fwd_f32(b2q_ctx* ctx, b2q_multi_plan* plan, int fix_alpha, int do_round, float qlevel,
                                  float ktimes, void* stream);
int b2q_multi_weight_ste_bwd_f32(b2q_ctx* ctx, b2q_multi_plan* plan, void* stream);

/* ---- cross-rank threshold exchange over peer memory (NVLink / NVSwitch), fused into the forward kernels ----
 * Data-parallel training needs allreduce(max) of every activation node's statistic before its threshold update
 * (BASELINE.json north_star).  Instead of reduce kernel -> ncclAllReduce(4 bytes) -> update kernel -> QDQ kernel,
 * b2q_peer_minmax_quant_fwd_f32 runs two kernels: the reduction's last block stores (sequence, max|x|) into every
 * rank's mailbox with 8-byte P2P stores, and the QDQ sweep reads its own mailbox (the first warps on each SM wait
 * until all `world` entries of that sequence number are present, take their max and cache it for the SM's later
 * blocks), applies the EMA / first-batch update in registers and sweeps.
 * Semantics = Quantization_int8 / ClipGrad_Quantization_int8 activation forward in training mode
 * (symbol/quant_ops.py:32-40, symbol/clip_grad_quantization_int8.py:37-51) with max|x| taken over all ranks.
 * mailboxes[r] = rank r's mailbox as mapped on THIS device (own: b2q_peer_mailbox_create; peers: the 64-byte CUDA
 * IPC handle exchanged out of band and opened with b2q_peer_mailbox_open).  All ranks must issue the same sequence
 * of calls; the sequence number itself is kept on the device (so a CUDA graph can replay the pair), the `sequence`
 * argument is informational.  A peer that never arrives makes the kernel trap after ~20 s instead of hanging.   */
int b2q_peer_mailbox_bytes(void);
/* The wait for a peer's statistic is bounded by time (option "peer_timeout_ms", default 600 000): when it expires the
 * sweep uses NaN as the statistic (threshold and output turn NaN), records the sequence number and the first missing
 * rank in its mailbox and carries on -- the CUDA context stays usable.  b2q_peer_status reads that record (0 = no
 * timeout so far); it copies device to host and therefore synchronises with the device.                            */
int b2q_peer_status(b2q_ctx* ctx, const void* own_mailbox, uint32_t* timeout_sequence, uint32_t* timeout_rank);
int b2q_peer_mailbox_create(b2q_ctx* ctx, void** mailbox, void* ipc_handle_out /* 64 bytes */);
int b2q_peer_mailbox_open(b2q_ctx* ctx, const void* ipc_handle /* 64 bytes */, void** peer_ptr);
int b2q_peer_mailbox_close(b2q_ctx* ctx, void* peer_ptr);
int b2q_peer_mailbox_destroy(b2q_ctx* ctx, void* mailbox);
/* A device buffer other ranks can map (cudaMalloc + CUDA IPC handle, zero-filled); map / unmap / free with
 * b2q_peer_mailbox_open / _close / _destroy.  For the gradient bucket and the statistic vectors below.            */
int b2q_peer_buffer_create(b2q_ctx* ctx, int64_t bytes, void** buffer, void* ipc_handle_out /* 64 bytes */);
/* Allreduce of one float32 buffer per rank over peer memory, one process per GPU, ONE kernel per rank and no NCCL:
 * gradients (sum, KVStore 'device' semantics of core/solver.py:121; average != 0 divides by world) and per-group
 * threshold statistics (max; grouped GDRQ_PY activations, core/operator/GDRQ.py:88-118 under data parallelism).
 * bufs[r] = rank r's buffer as mapped on THIS device, mailboxes as for the threshold exchange.  Rank r reduces the
 * r-th slice of every rank's buffer in rank order and stores it into every rank's buffer (bit-identical everywhere);
 * two flag barriers through the mailboxes order it with the peers' streams.  All ranks call it in the same order
 * with the same count; asynchronous on `stream`; waits bounded by "peer_timeout_ms" (b2q_peer_status).  The calls of
 * one set of mailboxes share one sequence counter: issue them stream-ordered with each other (not concurrently on two
 * streams).                                                                                                   */
int b2q_peer_allreduce_sum_f32(b2q_ctx* ctx, float* const* bufs, int64_t count, int average, void* const* mailboxes,
                               int rank, int world, void* stream);
int b2q_peer_allreduce_max_f32(b2q_ctx* ctx, float* const* bufs, int64_t count, void* const* mailboxes, int rank, int world,
                               void* stream);
int b2q_peer_minmax_quant_fwd_f32(b2q_ctx* ctx, int variant, const float* x, float* y, float* aux, int64_t n,
                                  int init, float ema_decay, float one_minus_decay, void* const* mailboxes,
                                  int rank, int world, uint32_t sequence, void* stream);
/* The same exchange for the operators whose threshold comes from mean|x| (statistic = max over ranks of the per-rank
 * mean): whole-tensor activations of GDRQ_PY (upd_mode B2Q_UPD_GDRQ_ACT, p0 = ktimes,
 * p1 = lamda, qlevel = 2^nbits-1; core/operator/GDRQ.py:69-86 with rounding on) and the data path of GDRQ_Fold_BN
 * (B2Q_UPD_TWICE_STORE on the first batch, then B2Q_UPD_TWICE_EMA, p0 = ema_decay, p1 = 1-ema_decay, qlevel = 127:
 * clip with the batch threshold, scale with the EMA one; symbol/fold_bn_v1_gdrq.py:53-68).                       */
int b2q_peer_meanabs_quant_fwd_f32(b2q_ctx* ctx, int upd_mode, const float* x, float* y, float* aux, int64_t n,
                                   float p0, float p1, float qlevel, void* const* mailboxes, int rank, int world,
                                   void* stream);

/* ---- one process, N devices (the reference's layout: train.py:34, core/solver.py:58-61) ---------------------------
 * The same two exchanges without torch.distributed, NCCL or CUDA IPC.  b2q_comm_create takes one context per device,
 * enables peer access between them and allocates a mailbox per rank; b2q_comm_mailboxes(comm, r, &boxes) yields the
 * table rank r passes to b2q_peer_minmax_quant_fwd_f32 / b2q_peer_meanabs_quant_fwd_f32 (world = b2q_comm_size), so
 * the fused threshold exchange works unchanged.  b2q_comm_allreduce_{max,sum}_f32 reduce one float32 buffer per rank
 * in place (bufs[r] on rank r's device, streams[r] its stream): rank r's kernel reads slice r from every rank over
 * NVLink, combines in rank order and stores the result into every rank's buffer -- bit-identical results everywhere,
 * stream-ordered across devices by events, asynchronous.  sum: KVStore 'device' semantics (solver.py:121); average != 0
 * divides by the rank count.                                                                                  */
typedef struct b2q_comm b2q_comm;
int b2q_comm_create(b2q_ctx* const* ctxs, int n, b2q_comm** out);
int b2q_comm_destroy(b2q_comm* comm);
int b2q_comm_size(b2q_comm* comm);
int b2q_comm_mailboxes(b2q_comm* comm, int rank, void* const** mailboxes);
int b2q_comm_allreduce_max_f32(b2q_comm* comm, float* const* bufs, int64_t count, void* const* streams);
int b2q_comm_allreduce_sum_f32(b2q_comm* comm, float* const* bufs, int64_t count, int average, void* const* streams);

/* ---- host-buffer path: the call a framework whose tensors live in HOST memory makes (bench.py "e2e") --
 * Same semantics as the device entry points but x / y / aux are HOST pointers (pinned for full speed).
 * Each call stages its tensor through a device staging ring on three internal streams
 * (H2D -> reduction + threshold update -> QDQ sweep -> D2H, per-segment events) and returns without waiting, so
 * the H2D of the next call, the kernels of this one and the D2H of the previous one overlap; b2q_host_sync()
 * waits for all of them.  host_aux is read at enqueue
 * order and written back by the same call; calls sharing an aux array must be separated by a sync.       */
int b2q_minmax_quant_fwd_host_f32(b2q_ctx* ctx, int variant, const float* host_x, float* host_y,
                                  float* host_aux, int64_t rows, int64_t cols, int is_weight,
                                  int per_channel, int is_train, int init, float ema_decay,
                                  float one_minus_decay);
int b2q_ste_bwd_host_f32(b2q_ctx* ctx, const float* host_dy, float* host_dx, int64_t n);
int b2q_clipgrad_bwd_host_f32(b2q_ctx* ctx, const float* host_x, const float* host_dy, float* host_dx,
                              const float* host_aux, int64_t n);
int b2q_host_sync(b2q_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* B2Q_H_ */
