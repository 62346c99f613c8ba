// libb2q.so -- lifecycle, primitives and the first-tier fused operator entry points (include/b2q.h).
#include <cstring>
#include <string>

#include "b2q_common.cuh"
#include "b2q_qdq.cuh"
#include "b2q_reduce.cuh"
#include "b2q_resident.cuh"
#include "b2q_cluster.cuh"

static thread_local std::string g_last_error;

void b2q_set_error(const std::string& msg) { g_last_error = msg; }

#define B2Q_CTX(ctx)                                                       \
    B2Q_REQUIRE((ctx) != nullptr, "null context");                         \
    B2Q_CHECK_CUDA(cudaSetDevice((ctx)->device))

static const Prescale kNoPrescale = {nullptr, nullptr, 0.f};
static const FoldBias kNoBias = {nullptr, nullptr, nullptr};

extern "C" {

int b2q_abi_version(void) { return B2Q_ABI_VERSION; }

const char* b2q_last_error(void) { return g_last_error.c_str(); }

int b2q_create(int device, b2q_ctx** out) {
    B2Q_REQUIRE(out != nullptr, "null out pointer");
    int count = 0;
    B2Q_CHECK_CUDA(cudaGetDeviceCount(&count));
    B2Q_REQUIRE(device >= 0 && device < count, "no such CUDA device");
    B2Q_CHECK_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    B2Q_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
    b2q_ctx* ctx = new b2q_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    cudaError_t e = cudaMalloc(&ctx->slots, sizeof(b2q_slot) * B2Q_NSLOTS);
    if (e != cudaSuccess) {
        delete ctx;
        b2q_set_error(std::string("cudaMalloc(workspace) failed: ") + cudaGetErrorString(e));
        return 1;
    }
    e = cudaMemset(ctx->slots, 0, sizeof(b2q_slot) * B2Q_NSLOTS);
    if (e != cudaSuccess) {
        cudaFree(ctx->slots);
        delete ctx;
        b2q_set_error(std::string("workspace init failed: ") + cudaGetErrorString(e));
        return 1;
    }
    *out = ctx;
    return 0;
}

int b2q_destroy(b2q_ctx* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    b2q_host_release(ctx);
    for (const b2q_timing_rec& r : ctx->recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
    cudaFree(ctx->slots);
    delete ctx;
    return 0;
}

int b2q_num_sms(b2q_ctx* ctx) { return ctx ? ctx->num_sms : -1; }

int b2q_stream_synchronize(b2q_ctx* ctx, void* stream) {
    B2Q_CTX(ctx);
    B2Q_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

int64_t b2q_launch_count(b2q_ctx* ctx) { return ctx ? ctx->launches : -1; }

static int* option_slot(b2q_ctx* ctx, const char* key) {
    if (!strcmp(key, "blocks_per_sm")) return &ctx->blocks_per_sm;
    if (!strcmp(key, "reduce_blocks_per_sm")) return &ctx->reduce_blocks_per_sm;
    if (!strcmp(key, "peer_reduce_blocks_per_sm")) return &ctx->peer_reduce_blocks_per_sm;
    if (!strcmp(key, "reduce_deferred_blocks_per_sm")) return &ctx->reduce_deferred_blocks_per_sm;
    if (!strcmp(key, "deferred")) return &ctx->deferred;
    if (!strcmp(key, "pdl")) return &ctx->pdl;
    if (!strcmp(key, "reverse")) return &ctx->reverse;
    if (!strcmp(key, "reverse_min_mb")) return &ctx->reverse_min_mb;
    if (!strcmp(key, "fast_div")) return &ctx->fast_div;
    if (!strcmp(key, "timing")) return &ctx->timing;
    if (!strcmp(key, "dorefa_tanh_max")) return &ctx->dorefa_tanh_max;
    if (!strcmp(key, "host_ste_copy")) return &ctx->host_ste_copy;
    if (!strcmp(key, "shared_slot_rings")) return &ctx->shared_rings;
    if (!strcmp(key, "cluster_fwd")) return &ctx->cluster_fwd;
    if (!strcmp(key, "cluster_max_elems")) return &ctx->cluster_max_elems;
    if (!strcmp(key, "cluster_max_elems_mean")) return &ctx->cluster_max_elems_mean;
    if (!strcmp(key, "cluster_words_per_cta")) return &ctx->cluster_words_per_cta;
    if (!strcmp(key, "resident")) return &ctx->resident;
    if (!strcmp(key, "resident_max_mb")) return &ctx->resident_max_mb;
    if (!strcmp(key, "peer_mode")) return &ctx->peer_mode;
    if (!strcmp(key, "seg_masked")) return &ctx->seg_masked;
    if (!strcmp(key, "stream_reduce")) return &ctx->stream_reduce;
    if (!strcmp(key, "stream_stages")) return &ctx->stream_stages;
    if (!strcmp(key, "stream_icvt")) return &ctx->stream_icvt;
    if (!strcmp(key, "bn_variant")) return &ctx->bn_variant;
    if (!strcmp(key, "bn_pieces_per_sm")) return &ctx->bn_pieces_per_sm;
    if (!strcmp(key, "peer_allreduce_blocks")) return &ctx->peer_allreduce_blocks;
    if (!strcmp(key, "peer_allreduce_blocks_per_sm")) return &ctx->peer_allreduce_blocks_per_sm;
    if (!strcmp(key, "peer_publish_blocks_per_sm")) return &ctx->peer_publish_blocks_per_sm;
    if (!strcmp(key, "peer_stage_early")) return &ctx->peer_stage_early;
    if (!strcmp(key, "peer_timeout_ms")) return &ctx->peer_timeout_ms;
    return nullptr;
}

int b2q_set_option(b2q_ctx* ctx, const char* key, int value) {
    B2Q_REQUIRE(ctx && key, "null argument");
    int* p = option_slot(ctx, key);
    B2Q_REQUIRE(p != nullptr, "unknown option");
    if (p == &ctx->blocks_per_sm || p == &ctx->reduce_blocks_per_sm || p == &ctx->peer_reduce_blocks_per_sm ||
        p == &ctx->reduce_deferred_blocks_per_sm)
        B2Q_REQUIRE(value >= 1 && value <= 65536, "blocks_per_sm out of range");
    if (p == &ctx->reduce_blocks_per_sm) B2Q_REQUIRE(value <= B2Q_MAX_PIECES / 256, "too many reduction blocks");
    *p = value;
    return 0;
}

int b2q_timing_read_range(b2q_ctx* ctx, int kind, double min_bytes, double max_bytes, double* total_ms,
                          double* total_bytes, int64_t* count, int reset) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(kind >= 0 && kind < B2Q_NKINDS, "bad kind");
    double ms = 0.0, bytes = 0.0;
    int64_t n = 0;
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (const b2q_timing_rec& r : ctx->recs) {
        if (kind != 0 && r.kind != kind) continue;
        if (r.bytes < min_bytes || (max_bytes > 0.0 && r.bytes >= max_bytes)) continue;
        B2Q_CHECK_CUDA(cudaEventSynchronize(r.e1));
        float t = 0.f;
        B2Q_CHECK_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
        ms += t; bytes += r.bytes; ++n;
    }
    if (total_ms) *total_ms = ms;
    if (total_bytes) *total_bytes = bytes;
    if (count) *count = n;
    if (reset) {
        for (const b2q_timing_rec& r : ctx->recs) { ctx->event_pool.push_back(r.e0); ctx->event_pool.push_back(r.e1); }
        ctx->recs.clear();
    }
    return 0;
}

int b2q_timing_read(b2q_ctx* ctx, int kind, double* total_ms, double* total_bytes, int64_t* count, int reset) {
    return b2q_timing_read_range(ctx, kind, 0.0, 0.0, total_ms, total_bytes, count, reset);
}

int b2q_get_option(b2q_ctx* ctx, const char* key, int* value) {
    B2Q_REQUIRE(ctx && key && value, "null argument");
    int* p = option_slot(ctx, key);
    B2Q_REQUIRE(p != nullptr, "unknown option");
    *value = *p;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// primitives
// ------------------------------------------------------------------------------------------------
static UpdateArgs stat_only(float* stat) {
    UpdateArgs u;
    memset(&u, 0, sizeof(u));
    u.stat_out = stat;
    return u;
}

int b2q_absmax_f32(b2q_ctx* ctx, const float* x, int64_t outer, int64_t groups, int64_t inner, float* stat,
                   void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(x && stat, "null pointer");
    return launch_reduce<true>(ctx, b2q_take_slot(ctx, (cudaStream_t)stream), x, outer, groups, inner, kNoPrescale, stat_only(stat),
                               (cudaStream_t)stream);
}

int b2q_meanabs_f32(b2q_ctx* ctx, const float* x, int64_t outer, int64_t groups, int64_t inner, float* stat,
                    void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(x && stat, "null pointer");
    return launch_reduce<false>(ctx, b2q_take_slot(ctx, (cudaStream_t)stream), x, outer, groups, inner, kNoPrescale, stat_only(stat),
                                (cudaStream_t)stream);
}

int b2q_threshold_update_f32(b2q_ctx* ctx, int mode, const float* stat, float* aux, float* clip_out,
                             int64_t groups, float p0, float p1, void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(stat && aux, "null pointer");
    B2Q_REQUIRE(mode >= B2Q_UPD_STORE && mode <= B2Q_UPD_TWICE_EMA, "bad update mode");
    B2Q_REQUIRE(groups >= 1 && groups <= B2Q_MAX_GROUPS, "bad group count");
    UpdateArgs u;
    memset(&u, 0, sizeof(u));
    u.mode = mode; u.write_aux = 1; u.use_aux_as_scale = 1; u.p0 = p0; u.p1 = p1; u.aux = aux; u.clip_out = clip_out;
    b2q_launch(ctx, threshold_update_kernel, (unsigned)((groups + 127) / 128), 128, (cudaStream_t)stream, stat, (int)groups, u);
    B2Q_LAUNCH_CHECK(ctx);
    return 0;
}

int b2q_qdq_f32(b2q_ctx* ctx, const float* x, float* y, int64_t outer, int64_t groups, int64_t inner,
                const float* thr, const float* clip_thr, float qlevel, int clip_mode, int do_round, int req,
                int32_t* codes, const float* prescale_gamma, const float* prescale_var, float prescale_eps,
                void* stream) {
    B2Q_CTX(ctx);
    if (req == B2Q_REQ_NULL) return 0;
    B2Q_REQUIRE(x && y && thr, "null pointer");
    QdqArgs a = {thr, clip_thr, 0.f, 0.f, qlevel, ctx->fast_div, codes, clip_mode, do_round, req};
    Prescale ps = {prescale_gamma, prescale_var, prescale_eps};
    return launch_qdq(ctx, x, y, outer, groups, inner, ps, kNoBias, a, (cudaStream_t)stream);
}

int b2q_ste_bwd_f32(b2q_ctx* ctx, const float* dy, float* dx, int64_t n, int req, void* stream) {
    B2Q_CTX(ctx);
    if (req == B2Q_REQ_NULL) return 0;
    B2Q_REQUIRE(dy && dx && n >= 1, "bad argument");
    if (dy == dx && req != B2Q_REQ_ADD) return 0;  // in-place identity
    return launch_bwd_mask<0>(ctx, nullptr, dy, dx, 1, 1, n, nullptr, 0.f, req, (cudaStream_t)stream);
}

int b2q_zero_f32(b2q_ctx* ctx, float* x, int64_t n, void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(x != nullptr && n >= 0, "bad argument");
    if (n == 0) return 0;
    B2Q_CHECK_CUDA(cudaMemsetAsync(x, 0, sizeof(float) * (size_t)n, (cudaStream_t)stream));
    return 0;
}

int b2q_mask_bwd_f32(b2q_ctx* ctx, const float* x, const float* dy, float* dx, int64_t outer, int64_t groups,
                     int64_t inner, const float* thr, float thr_imm, int mask_mode, int req, void* stream) {
    B2Q_CTX(ctx);
    if (req == B2Q_REQ_NULL) return 0;
    B2Q_REQUIRE(x && dy && dx, "null pointer");
    B2Q_REQUIRE(outer >= 1 && groups >= 1 && inner >= 1, "empty tensor");
    cudaStream_t st = (cudaStream_t)stream;
    switch (mask_mode) {
        case B2Q_MASK_OPEN: return launch_bwd_mask<B2Q_MASK_OPEN>(ctx, x, dy, dx, outer, groups, inner, thr, thr_imm, req, st);
        case B2Q_MASK_ABS_LE: return launch_bwd_mask<B2Q_MASK_ABS_LE>(ctx, x, dy, dx, outer, groups, inner, thr, thr_imm, req, st);
        case B2Q_MASK_LT: return launch_bwd_mask<B2Q_MASK_LT>(ctx, x, dy, dx, outer, groups, inner, thr, thr_imm, req, st);
        default: break;
    }
    b2q_set_error("b2q: unknown mask_mode");
    return 2;
}

// ------------------------------------------------------------------------------------------------
// Quantization_int8_V2 / ClipGrad_Quantization_int8
// ------------------------------------------------------------------------------------------------
static int minmax_groups(int64_t rows, int64_t cols, int per_channel, int64_t* outer, int64_t* groups, int64_t* inner) {
    *outer = 1;
    if (per_channel) { *groups = rows; *inner = cols; }
    else { *groups = 1; *inner = rows * cols; }
    return 0;
}

// Threshold bookkeeping of the two minmax operators.  Returns which buffer the sweep scales with.
static UpdateArgs minmax_update(int variant, int is_weight, int is_train, int init, float d, float omd, float* aux,
                                b2q_slot* slot, const float** scale_src) {
    UpdateArgs u;
    memset(&u, 0, sizeof(u));
    u.p0 = d; u.p1 = omd; u.aux = aux;
    if (is_weight) {
        u.mode = B2Q_UPD_STORE;
        u.write_aux = is_train ? 1 : 0;
        if (variant == 0) {               // quant_ops.py:26-31: scale with the fresh max, store it when training
            u.use_aux_as_scale = 0;
            u.scale_out = slot->scale;
            *scale_src = slot->scale;
        } else {                          // clip_grad...py:31-35: aux is the source of truth
            u.use_aux_as_scale = 1;
            *scale_src = aux;
        }
    } else {
        u.mode = (variant == 1 && init) ? B2Q_UPD_STORE : B2Q_UPD_EMA;   // clip_grad...py:42-46 / quant_ops.py:37
        u.write_aux = 1;
        u.use_aux_as_scale = 1;
        *scale_src = aux;
    }
    return u;
}

int b2q_minmax_quant_fwd_f32(b2q_ctx* ctx, int variant, const float* x, float* y, float* aux, int64_t rows,
                             int64_t cols, int is_weight, int per_channel, int is_train, int init, float ema_decay,
                             float one_minus_decay, int req, void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(x && y && aux, "null pointer");
    B2Q_REQUIRE(variant == 0 || variant == 1, "variant must be 0 (Quantization_int8_V2) or 1 (ClipGrad)");
    B2Q_REQUIRE(rows >= 1 && cols >= 1, "empty tensor");
    B2Q_REQUIRE(!(per_channel && !is_weight), "per-channel thresholds are a weight-only feature");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t outer, groups, inner;
    minmax_groups(rows, cols, per_channel && is_weight, &outer, &groups, &inner);
    const float* scale_src = aux;
    // Does this call need a reduction?  weight: V2 always, ClipGrad only when training; act: when training.
    const bool reduce = is_weight ? (variant == 0 || is_train) : (is_train != 0);
    // ClipGrad activations are clipped to +-aux and written with [:]= regardless of req (clip_grad...py:48-51)
    const bool clip = (variant == 1 && !is_weight);
    int eff_req = req;
    if (clip) eff_req = B2Q_REQ_WRITE;
    if (reduce) {
        b2q_slot* slot = b2q_take_slot(ctx, st);
        UpdateArgs u = minmax_update(variant, is_weight, is_train, init, ema_decay, one_minus_decay, aux, slot, &scale_src);
        if (groups == 1 && (eff_req == B2Q_REQ_WRITE || eff_req == B2Q_REQ_INPLACE)) {
            // fused: partial maxima + a sweep that finishes the threshold update itself (no serialized tail)
            UpdateArgs ud = u;
            ud.scale_out = nullptr;
            int done = 0;
            // small tensors (weights): one launch of one thread-block cluster, the tensor read once (b2q_cluster.cuh)
            int rc = launch_fused_cluster<true>(ctx, x, y, inner, ud, 127.f, clip ? B2Q_CLIP_SYM : B2Q_CLIP_NONE, 0, st, &done);
            if (rc || done) return rc;
            // tensors that fit on chip: one launch (reduce -> grid barrier -> sweep out of shared memory / L2)
            rc = launch_fused_resident<true>(ctx, slot, x, y, inner, ud, 127.f, clip ? B2Q_CLIP_SYM : B2Q_CLIP_NONE, 0,
                                                 st, &done);
            if (rc || done) return rc;
            rc = launch_fused_flat_fwd<true>(ctx, slot, x, y, inner, ud, 127.f, clip ? B2Q_CLIP_SYM : B2Q_CLIP_NONE, 0,
                                             st, &done);
            if (rc || done) return rc;
        }
        if (groups > 1) {   // per-channel weight: one warp per row does reduce + update + sweep in one launch
            UpdateArgs ur = u;
            ur.scale_out = nullptr;
            QdqArgs ar = {nullptr, nullptr, 0.f, 0.f, 127.f, ctx->fast_div, nullptr, B2Q_CLIP_NONE, 1, eff_req};
            int done = 0;
            int rc = launch_rows_fused<true>(ctx, x, y, groups, inner, kNoPrescale, kNoBias, ur, ar, 0, st, &done);
            if (rc || done) return rc;
        }
        int rc = launch_reduce<true>(ctx, slot, x, outer, groups, inner, kNoPrescale, u, st);
        if (rc) return rc;
    }
    if (eff_req == B2Q_REQ_NULL) return 0;
    QdqArgs a = {scale_src, nullptr, 0.f, 0.f, 127.f, ctx->fast_div, nullptr,
                 clip ? B2Q_CLIP_SYM : B2Q_CLIP_NONE, 1, eff_req};
    return launch_qdq(ctx, x, y, outer, groups, inner, kNoPrescale, kNoBias, a, st);
}

int b2q_minmax_quant_stat_f32(b2q_ctx* ctx, const float* x, int64_t rows, int64_t cols, int per_channel,
                              float* stat_out, void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(x && stat_out, "null pointer");
    int64_t outer, groups, inner;
    minmax_groups(rows, cols, per_channel, &outer, &groups, &inner);
    return launch_reduce<true>(ctx, b2q_take_slot(ctx, (cudaStream_t)stream), x, outer, groups, inner, kNoPrescale, stat_only(stat_out),
                               (cudaStream_t)stream);
}

int b2q_minmax_quant_finish_f32(b2q_ctx* ctx, int variant, const float* x, float* y, float* aux, const float* stat,
                                int64_t rows, int64_t cols, int is_weight, int per_channel, int is_train, int init,
                                float ema_decay, float one_minus_decay, int req, void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(x && y && aux && stat, "null pointer");
    B2Q_REQUIRE(variant == 0 || variant == 1, "variant must be 0 or 1");
    cudaStream_t st = (cudaStream_t)stream;
    int64_t outer, groups, inner;
    minmax_groups(rows, cols, per_channel && is_weight, &outer, &groups, &inner);
    B2Q_REQUIRE(groups <= B2Q_MAX_GROUPS, "too many channels");
    b2q_slot* slot = b2q_take_slot(ctx, st);
    const float* scale_src = aux;
    UpdateArgs u = minmax_update(variant, is_weight, is_train, init, ema_decay, one_minus_decay, aux, slot, &scale_src);
    b2q_launch(ctx, threshold_update_kernel, (unsigned)((groups + 127) / 128), 128, st, stat, (int)groups, u);
    B2Q_LAUNCH_CHECK(ctx);
    const bool clip = (variant == 1 && !is_weight);
    int eff_req = clip ? B2Q_REQ_WRITE : req;
    if (eff_req == B2Q_REQ_NULL) return 0;
    QdqArgs a = {scale_src, nullptr, 0.f, 0.f, 127.f, ctx->fast_div, nullptr,
                 clip ? B2Q_CLIP_SYM : B2Q_CLIP_NONE, 1, eff_req};
    return launch_qdq(ctx, x, y, outer, groups, inner, kNoPrescale, kNoBias, a, st);
}

int b2q_clipgrad_bwd_f32(b2q_ctx* ctx, const float* x, const float* dy, float* dx, const float* aux, int64_t n,
                         void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(x && dy && dx && aux && n >= 1, "bad argument");
    // written with [:]= in the reference (clip_grad...py:61-67): req is ignored
    return launch_bwd_mask<B2Q_MASK_OPEN>(ctx, x, dy, dx, 1, 1, n, aux, 0.f, B2Q_REQ_WRITE, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// GDRQ_PY
// ------------------------------------------------------------------------------------------------
int b2q_gdrq_fwd_f32(b2q_ctx* ctx, const float* x, float* y, float* alpha, int64_t outer, int64_t groups,
                     int64_t inner, int is_weight, int fix_alpha, int do_round, float qlevel, float ktimes,
                     float lamda, int req, void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(x && y && alpha, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (!fix_alpha) {  // GDRQ.py:70-76 / :97-104 -- runs in eval mode too
        UpdateArgs u;
        memset(&u, 0, sizeof(u));
        u.mode = is_weight ? B2Q_UPD_GDRQ_WEIGHT : B2Q_UPD_GDRQ_ACT;
        u.write_aux = 1; u.use_aux_as_scale = 1; u.p0 = ktimes; u.p1 = lamda; u.aux = alpha;
        b2q_slot* slot = b2q_take_slot(ctx, st);
        if (groups == 1 && do_round && (req == B2Q_REQ_WRITE || req == B2Q_REQ_INPLACE)) {
            int done = 0;
            int rc = launch_fused_cluster<false>(ctx, x, y, outer * inner, u, qlevel, B2Q_CLIP_SYM, 0, st, &done);
            if (rc || done) return rc;
            rc = launch_fused_resident<false>(ctx, slot, x, y, outer * inner, u, qlevel, B2Q_CLIP_SYM, 0, st, &done);
            if (rc || done) return rc;
            rc = launch_fused_flat_fwd<false>(ctx, slot, x, y, outer * inner, u, qlevel, B2Q_CLIP_SYM, 0, st, &done);
            if (rc || done) return rc;
        }
        if (groups > 1 && outer == 1) {   // grouped weight: warp per group, one launch
            QdqArgs ar = {nullptr, nullptr, 0.f, 0.f, qlevel, ctx->fast_div, nullptr, B2Q_CLIP_WHERE_LE, do_round, req};
            int done = 0;
            int rc = launch_rows_fused<false>(ctx, x, y, groups, inner, kNoPrescale, kNoBias, u, ar, 0, st, &done);
            if (rc || done) return rc;
        }
        int rc = launch_reduce<false>(ctx, slot, x, outer, groups, inner, kNoPrescale, u, st);
        if (rc) return rc;
    }
    if (req == B2Q_REQ_NULL) return 0;
    // group_size == -1 clips with mx.nd.clip (GDRQ.py:79); grouped mode with where(|x|<=a, x, a*sign(x)) (:109)
    const int clip = (groups == 1) ? B2Q_CLIP_SYM : B2Q_CLIP_WHERE_LE;
    QdqArgs a = {alpha, nullptr, 0.f, 0.f, qlevel, ctx->fast_div, nullptr, clip, do_round, req};
    return launch_qdq(ctx, x, y, outer, groups, inner, kNoPrescale, kNoBias, a, st);
}

int b2q_gdrq_bwd_f32(b2q_ctx* ctx, const float* x, const float* dy, float* dx, const float* alpha, int64_t outer,
                     int64_t groups, int64_t inner, int req, void* stream) {
    B2Q_CTX(ctx);
    if (req == B2Q_REQ_NULL) return 0;
    B2Q_REQUIRE(x && dy && dx && alpha, "null pointer");
    return launch_bwd_mask<B2Q_MASK_ABS_LE>(ctx, x, dy, dx, outer, groups, inner, alpha, 0.f, req, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------
// GDRQ_Fold_BN
// ------------------------------------------------------------------------------------------------
int b2q_foldbn_data_fwd_f32(b2q_ctx* ctx, const float* x, float* y, float* aux_data, int64_t n, int init,
                            float ema_decay, float one_minus_decay, void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(x && y && aux_data && n >= 1, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    b2q_slot* slot = b2q_take_slot(ctx, st);
    UpdateArgs u;
    memset(&u, 0, sizeof(u));
    u.mode = init ? B2Q_UPD_TWICE_STORE : B2Q_UPD_TWICE_EMA;   // fold_bn_v1_gdrq.py:58-64
    u.write_aux = 1; u.use_aux_as_scale = 1; u.p0 = ema_decay; u.p1 = one_minus_decay; u.aux = aux_data;
    {
        int done = 0;
        int rc = launch_fused_cluster<false>(ctx, x, y, n, u, 127.f, B2Q_CLIP_SYM, /*clip_with_fresh=*/1, st, &done);
        if (rc || done) return rc;
        rc = launch_fused_resident<false>(ctx, slot, x, y, n, u, 127.f, B2Q_CLIP_SYM, /*clip_with_fresh=*/1, st, &done);
        if (rc || done) return rc;
        rc = launch_fused_flat_fwd<false>(ctx, slot, x, y, n, u, 127.f, B2Q_CLIP_SYM, /*clip_with_fresh=*/1, st, &done);
        if (rc || done) return rc;
    }
    u.clip_out = slot->clip;                                    // :67 clips with the batch threshold
    int rc = launch_reduce<false>(ctx, slot, x, 1, 1, n, kNoPrescale, u, st);
    if (rc) return rc;
    QdqArgs a = {aux_data, slot->clip, 0.f, 0.f, 127.f, ctx->fast_div, nullptr, B2Q_CLIP_SYM, 1, B2Q_REQ_WRITE};
    return launch_qdq(ctx, x, y, 1, 1, n, kNoPrescale, kNoBias, a, st);
}

__global__ void foldbn_scale_only_kernel(const float* __restrict__ w, float* __restrict__ wq, float* __restrict__ bias,
                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                         const float* __restrict__ mean, const float* __restrict__ var, float eps,
                                         int64_t cols) {
    // quantize_flag == False: weight * factor and the folded bias only (fold_bn_v1_gdrq.py:70-74,113)
    const int64_t row = blockIdx.x;
    const float den = __fsqrt_rn(__fadd_rn(var[row], eps));
    const float f = __fdiv_rn(gamma[row], den);
    for (int64_t i = threadIdx.x; i < cols; i += blockDim.x) wq[row * cols + i] = __fmul_rn(w[row * cols + i], f);
    if (threadIdx.x == 0 && bias) bias[row] = __fsub_rn(beta[row], __fdiv_rn(__fmul_rn(mean[row], gamma[row]), den));
}

int b2q_foldbn_weight_fwd_f32(b2q_ctx* ctx, const float* w, float* w_q, float* bias, float* aux_weight,
                              const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                              int64_t cout, int64_t cols, int per_channel, int quantize, int is_train, void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(w && w_q && gamma && beta && mean && var, "null pointer");
    B2Q_REQUIRE(cout >= 1 && cols >= 1, "empty weight");
    cudaStream_t st = (cudaStream_t)stream;
    if (!quantize) {
        foldbn_scale_only_kernel<<<(unsigned)cout, 128, 0, st>>>(w, w_q, bias, gamma, beta, mean, var, eps, cols);
        B2Q_LAUNCH_CHECK(ctx);
        return 0;
    }
    B2Q_REQUIRE(aux_weight != nullptr, "null aux");
    b2q_slot* slot = b2q_take_slot(ctx, st);
    Prescale ps = {gamma, var, eps};
    FoldBias fb = {bias, beta, mean};
    // view: per-channel (1, cout, cols); per-tensor (cout, 1, cols) so that the prescale row is o*groups+g either way
    const int64_t outer = per_channel ? 1 : cout, groups = per_channel ? cout : 1;
    UpdateArgs u;
    memset(&u, 0, sizeof(u));
    u.mode = B2Q_UPD_TWICE_STORE;                 // fold_bn_v1_gdrq.py:82 / :90, aux stored only when training (:94-95)
    u.write_aux = is_train ? 1 : 0;
    u.use_aux_as_scale = 0;
    u.aux = aux_weight;
    if (per_channel) {   // warp per out-channel: prescale, 2*mean|w'|, clip, QDQ and the folded bias in one launch
        QdqArgs ar = {nullptr, nullptr, 0.f, 0.f, 127.f, ctx->fast_div, nullptr, B2Q_CLIP_SYM, 1, B2Q_REQ_WRITE};
        int done = 0;
        int rc = launch_rows_fused<false>(ctx, w, w_q, cout, cols, ps, fb, u, ar, 0, st, &done);
        if (rc || done) return rc;
    }
    u.scale_out = slot->scale;
    int rc = launch_reduce<false>(ctx, slot, w, outer, groups, cols, ps, u, st);
    if (rc) return rc;
    QdqArgs a = {slot->scale, nullptr, 0.f, 0.f, 127.f, ctx->fast_div, nullptr, B2Q_CLIP_SYM, 1, B2Q_REQ_WRITE};
    return launch_qdq(ctx, w, w_q, outer, groups, cols, ps, fb, a, st);
}

}  // extern "C"
