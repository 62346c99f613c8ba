// Host-buffer entry points: the call a framework makes when its tensors live in host memory
// (bench.py's "e2e" leg).
//
// A three-stream software pipeline over a ring of device staging sets keeps both PCIe directions busy:
//     h2d stream   : host x (and dy) -> set.a / set.b              continuous host-to-device traffic
//     compute strm : reduction + threshold update + QDQ sweep / backward mask on the staged tensors
//     d2h stream   : set output -> host y / dx, aux -> host       continuous device-to-host traffic
// Call k uses set k % B2Q_HOST_SETS; events chain h2d -> compute -> d2h within a set and d2h(k) -> h2d(k + SETS)
// across reuses.  Every call returns after enqueueing; b2q_host_sync() waits for all of them.  Calls that touch the
// same host aux array must be separated by b2q_host_sync().
#include <cstring>

#include "b2q_common.cuh"

#define B2Q_CTX(ctx)                               \
    B2Q_REQUIRE((ctx) != nullptr, "null context"); \
    B2Q_CHECK_CUDA(cudaSetDevice((ctx)->device))

#define B2Q_HOST_SETS 4

struct HostStage {
    float* a = nullptr;      // staged input
    float* b = nullptr;      // staged second input (dy) or output
    float* c = nullptr;      // output of two-input ops
    float* aux = nullptr;    // device mirror of the aux vector
    cudaEvent_t h2d_done = nullptr, comp_done = nullptr, d2h_done = nullptr;
    bool used = false;
};

struct HostState {
    HostStage set[B2Q_HOST_SETS];
    size_t cap = 0;          // elements per staging buffer
    bool have_c = false;
    unsigned next = 0;
    cudaStream_t h2d = nullptr, comp = nullptr, d2h = nullptr;
};

static HostState* host_state(b2q_ctx* ctx) { return reinterpret_cast<HostState*>(ctx->host_state); }

static int host_init(b2q_ctx* ctx) {
    if (ctx->host_state) return 0;
    HostState* hs = new HostState();
    ctx->host_state = hs;
    B2Q_CHECK_CUDA(cudaStreamCreateWithFlags(&hs->h2d, cudaStreamNonBlocking));
    B2Q_CHECK_CUDA(cudaStreamCreateWithFlags(&hs->comp, cudaStreamNonBlocking));
    B2Q_CHECK_CUDA(cudaStreamCreateWithFlags(&hs->d2h, cudaStreamNonBlocking));
    for (HostStage& s : hs->set) {
        B2Q_CHECK_CUDA(cudaMalloc(&s.aux, sizeof(float) * B2Q_MAX_GROUPS));
        B2Q_CHECK_CUDA(cudaEventCreateWithFlags(&s.h2d_done, cudaEventDisableTiming));
        B2Q_CHECK_CUDA(cudaEventCreateWithFlags(&s.comp_done, cudaEventDisableTiming));
        B2Q_CHECK_CUDA(cudaEventCreateWithFlags(&s.d2h_done, cudaEventDisableTiming));
    }
    return 0;
}

static int host_sync_all(HostState* hs) {
    B2Q_CHECK_CUDA(cudaStreamSynchronize(hs->h2d));
    B2Q_CHECK_CUDA(cudaStreamSynchronize(hs->comp));
    B2Q_CHECK_CUDA(cudaStreamSynchronize(hs->d2h));
    return 0;
}

// Next staging set with room for n elements (and a third buffer if need_c); the h2d stream is made to wait until the
// set's previous occupant has been copied out.
static int take_stage(b2q_ctx* ctx, int64_t n, bool need_c, HostStage** out) {
    int rc = host_init(ctx);
    if (rc) return rc;
    HostState* hs = host_state(ctx);
    if ((size_t)n > hs->cap || (need_c && !hs->have_c)) {
        rc = host_sync_all(hs);
        if (rc) return rc;
        size_t cap = (size_t)n > hs->cap ? (size_t)n : hs->cap;
        for (HostStage& s : hs->set) {
            if (cap > hs->cap) {
                cudaFree(s.a); cudaFree(s.b); cudaFree(s.c);
                s.a = s.b = s.c = nullptr;
                B2Q_CHECK_CUDA(cudaMalloc(&s.a, sizeof(float) * cap));
                B2Q_CHECK_CUDA(cudaMalloc(&s.b, sizeof(float) * cap));
            }
            if ((need_c || hs->have_c) && !s.c) B2Q_CHECK_CUDA(cudaMalloc(&s.c, sizeof(float) * cap));
        }
        hs->cap = cap;
        hs->have_c = hs->have_c || need_c;
    }
    HostStage& s = hs->set[hs->next++ % B2Q_HOST_SETS];
    if (s.used) B2Q_CHECK_CUDA(cudaStreamWaitEvent(hs->h2d, s.d2h_done, 0));
    s.used = true;
    *out = &s;
    return 0;
}

int b2q_host_release(b2q_ctx* ctx) {
    if (!ctx || !ctx->host_state) return 0;
    HostState* hs = host_state(ctx);
    if (hs->h2d) { cudaStreamSynchronize(hs->h2d); cudaStreamDestroy(hs->h2d); }
    if (hs->comp) { cudaStreamSynchronize(hs->comp); cudaStreamDestroy(hs->comp); }
    if (hs->d2h) { cudaStreamSynchronize(hs->d2h); cudaStreamDestroy(hs->d2h); }
    for (HostStage& s : hs->set) {
        cudaFree(s.a); cudaFree(s.b); cudaFree(s.c); cudaFree(s.aux);
        if (s.h2d_done) cudaEventDestroy(s.h2d_done);
        if (s.comp_done) cudaEventDestroy(s.comp_done);
        if (s.d2h_done) cudaEventDestroy(s.d2h_done);
    }
    delete hs;
    ctx->host_state = nullptr;
    return 0;
}

extern "C" {

int b2q_host_sync(b2q_ctx* ctx) {
    B2Q_CTX(ctx);
    if (!ctx->host_state) return 0;
    return host_sync_all(host_state(ctx));
}

int b2q_minmax_quant_fwd_host_f32(b2q_ctx* ctx, int variant, const float* host_x, float* host_y, float* host_aux,
                                  int64_t rows, int64_t cols, int is_weight, int per_channel, int is_train, int init,
                                  float ema_decay, float one_minus_decay) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(host_x && host_y && host_aux && rows >= 1 && cols >= 1, "bad argument");
    const int64_t n = rows * cols;
    const int64_t naux = (per_channel && is_weight) ? rows : 1;
    B2Q_REQUIRE(naux <= B2Q_MAX_GROUPS, "too many channels");
    HostStage* s;
    int rc = take_stage(ctx, n, false, &s);
    if (rc) return rc;
    HostState* hs = host_state(ctx);
    B2Q_CHECK_CUDA(cudaMemcpyAsync(s->aux, host_aux, sizeof(float) * naux, cudaMemcpyHostToDevice, hs->h2d));
    B2Q_CHECK_CUDA(cudaMemcpyAsync(s->a, host_x, sizeof(float) * n, cudaMemcpyHostToDevice, hs->h2d));
    B2Q_CHECK_CUDA(cudaEventRecord(s->h2d_done, hs->h2d));
    B2Q_CHECK_CUDA(cudaStreamWaitEvent(hs->comp, s->h2d_done, 0));
    rc = b2q_minmax_quant_fwd_f32(ctx, variant, s->a, s->b, s->aux, rows, cols, is_weight, per_channel, is_train, init,
                                  ema_decay, one_minus_decay, B2Q_REQ_WRITE, hs->comp);
    if (rc) return rc;
    B2Q_CHECK_CUDA(cudaEventRecord(s->comp_done, hs->comp));
    B2Q_CHECK_CUDA(cudaStreamWaitEvent(hs->d2h, s->comp_done, 0));
    B2Q_CHECK_CUDA(cudaMemcpyAsync(host_y, s->b, sizeof(float) * n, cudaMemcpyDeviceToHost, hs->d2h));
    B2Q_CHECK_CUDA(cudaMemcpyAsync(host_aux, s->aux, sizeof(float) * naux, cudaMemcpyDeviceToHost, hs->d2h));
    B2Q_CHECK_CUDA(cudaEventRecord(s->d2h_done, hs->d2h));
    return 0;
}

int b2q_ste_bwd_host_f32(b2q_ctx* ctx, const float* host_dy, float* host_dx, int64_t n) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(host_dy && host_dx && n >= 1, "bad argument");
    HostStage* s;
    int rc = take_stage(ctx, n, false, &s);
    if (rc) return rc;
    HostState* hs = host_state(ctx);
    B2Q_CHECK_CUDA(cudaMemcpyAsync(s->a, host_dy, sizeof(float) * n, cudaMemcpyHostToDevice, hs->h2d));
    B2Q_CHECK_CUDA(cudaEventRecord(s->h2d_done, hs->h2d));
    B2Q_CHECK_CUDA(cudaStreamWaitEvent(hs->comp, s->h2d_done, 0));
    rc = b2q_ste_bwd_f32(ctx, s->a, s->b, n, B2Q_REQ_WRITE, hs->comp);
    if (rc) return rc;
    B2Q_CHECK_CUDA(cudaEventRecord(s->comp_done, hs->comp));
    B2Q_CHECK_CUDA(cudaStreamWaitEvent(hs->d2h, s->comp_done, 0));
    B2Q_CHECK_CUDA(cudaMemcpyAsync(host_dx, s->b, sizeof(float) * n, cudaMemcpyDeviceToHost, hs->d2h));
    B2Q_CHECK_CUDA(cudaEventRecord(s->d2h_done, hs->d2h));
    return 0;
}

int b2q_clipgrad_bwd_host_f32(b2q_ctx* ctx, const float* host_x, const float* host_dy, float* host_dx,
                              const float* host_aux, int64_t n) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(host_x && host_dy && host_dx && host_aux && n >= 1, "bad argument");
    HostStage* s;
    int rc = take_stage(ctx, n, true, &s);
    if (rc) return rc;
    HostState* hs = host_state(ctx);
    B2Q_CHECK_CUDA(cudaMemcpyAsync(s->aux, host_aux, sizeof(float), cudaMemcpyHostToDevice, hs->h2d));
    B2Q_CHECK_CUDA(cudaMemcpyAsync(s->a, host_x, sizeof(float) * n, cudaMemcpyHostToDevice, hs->h2d));
    B2Q_CHECK_CUDA(cudaMemcpyAsync(s->b, host_dy, sizeof(float) * n, cudaMemcpyHostToDevice, hs->h2d));
    B2Q_CHECK_CUDA(cudaEventRecord(s->h2d_done, hs->h2d));
    B2Q_CHECK_CUDA(cudaStreamWaitEvent(hs->comp, s->h2d_done, 0));
    rc = b2q_clipgrad_bwd_f32(ctx, s->a, s->b, s->c, s->aux, n, hs->comp);
    if (rc) return rc;
    B2Q_CHECK_CUDA(cudaEventRecord(s->comp_done, hs->comp));
    B2Q_CHECK_CUDA(cudaStreamWaitEvent(hs->d2h, s->comp_done, 0));
    B2Q_CHECK_CUDA(cudaMemcpyAsync(host_dx, s->c, sizeof(float) * n, cudaMemcpyDeviceToHost, hs->d2h));
    B2Q_CHECK_CUDA(cudaEventRecord(s->d2h_done, hs->d2h));
    return 0;
}

}  // extern "C"
