// Host-buffer entry points: the call a framework makes when its tensors live in host memory
// (bench.py's "e2e" leg).  Each call stages its tensor through one of two device staging sets on that set's
// own stream: H2D copy -> reduction(+threshold update) -> QDQ sweep -> D2H copy, all asynchronous.  Because
// consecutive calls alternate between the two sets, the D2H of call k overlaps the H2D of call k+1 (PCIe is
// full duplex).  b2q_host_sync() waits for everything.  Calls that touch the same host aux array must be
// separated by b2q_host_sync().
#include <cstring>

#include "b2q_common.cuh"

#define B2Q_CTX(ctx)                               \
    B2Q_REQUIRE((ctx) != nullptr, "null context"); \
    B2Q_CHECK_CUDA(cudaSetDevice((ctx)->device))

struct HostStage {
    float* a = nullptr;      // device staging: input
    float* b = nullptr;      // device staging: second input (dy) / output
    float* c = nullptr;      // device staging: output of two-input ops
    float* aux = nullptr;    // device mirror of the aux vector
    size_t cap = 0;          // elements in a / b / c
    cudaStream_t stream = nullptr;
};

struct HostState {
    HostStage set[2];
    unsigned next = 0;
};

static HostState* host_state(b2q_ctx* ctx) { return reinterpret_cast<HostState*>(ctx->host_state); }

static int ensure_stage(b2q_ctx* ctx, int64_t n, bool need_c, HostStage** out) {
    if (!ctx->host_state) ctx->host_state = new HostState();
    HostState* hs = host_state(ctx);
    HostStage& s = hs->set[hs->next++ & 1];
    if (!s.stream) B2Q_CHECK_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    if (!s.aux) B2Q_CHECK_CUDA(cudaMalloc(&s.aux, sizeof(float) * B2Q_MAX_GROUPS));
    if ((size_t)n > s.cap) {
        B2Q_CHECK_CUDA(cudaStreamSynchronize(s.stream));
        cudaFree(s.a); cudaFree(s.b); cudaFree(s.c);
        s.a = s.b = s.c = nullptr;
        s.cap = 0;
        B2Q_CHECK_CUDA(cudaMalloc(&s.a, sizeof(float) * n));
        B2Q_CHECK_CUDA(cudaMalloc(&s.b, sizeof(float) * n));
        s.cap = (size_t)n;
    }
    if (need_c && !s.c) B2Q_CHECK_CUDA(cudaMalloc(&s.c, sizeof(float) * s.cap));
    *out = &s;
    return 0;
}

int b2q_host_release(b2q_ctx* ctx) {
    if (!ctx || !ctx->host_state) return 0;
    HostState* hs = host_state(ctx);
    for (HostStage& s : hs->set) {
        if (s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
        cudaFree(s.a); cudaFree(s.b); cudaFree(s.c); cudaFree(s.aux);
    }
    delete hs;
    ctx->host_state = nullptr;
    return 0;
}

extern "C" {

int b2q_host_sync(b2q_ctx* ctx) {
    B2Q_CTX(ctx);
    if (!ctx->host_state) return 0;
    HostState* hs = host_state(ctx);
    for (HostStage& s : hs->set)
        if (s.stream) B2Q_CHECK_CUDA(cudaStreamSynchronize(s.stream));
    return 0;
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
    int rc = ensure_stage(ctx, n, false, &s);
    if (rc) return rc;
    B2Q_CHECK_CUDA(cudaMemcpyAsync(s->aux, host_aux, sizeof(float) * naux, cudaMemcpyHostToDevice, s->stream));
    B2Q_CHECK_CUDA(cudaMemcpyAsync(s->a, host_x, sizeof(float) * n, cudaMemcpyHostToDevice, s->stream));
    rc = b2q_minmax_quant_fwd_f32(ctx, variant, s->a, s->b, s->aux, rows, cols, is_weight, per_channel, is_train, init,
                                  ema_decay, one_minus_decay, B2Q_REQ_WRITE, s->stream);
    if (rc) return rc;
    B2Q_CHECK_CUDA(cudaMemcpyAsync(host_y, s->b, sizeof(float) * n, cudaMemcpyDeviceToHost, s->stream));
    B2Q_CHECK_CUDA(cudaMemcpyAsync(host_aux, s->aux, sizeof(float) * naux, cudaMemcpyDeviceToHost, s->stream));
    return 0;
}

int b2q_ste_bwd_host_f32(b2q_ctx* ctx, const float* host_dy, float* host_dx, int64_t n) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(host_dy && host_dx && n >= 1, "bad argument");
    HostStage* s;
    int rc = ensure_stage(ctx, n, false, &s);
    if (rc) return rc;
    B2Q_CHECK_CUDA(cudaMemcpyAsync(s->a, host_dy, sizeof(float) * n, cudaMemcpyHostToDevice, s->stream));
    rc = b2q_ste_bwd_f32(ctx, s->a, s->b, n, B2Q_REQ_WRITE, s->stream);
    if (rc) return rc;
    B2Q_CHECK_CUDA(cudaMemcpyAsync(host_dx, s->b, sizeof(float) * n, cudaMemcpyDeviceToHost, s->stream));
    return 0;
}

int b2q_clipgrad_bwd_host_f32(b2q_ctx* ctx, const float* host_x, const float* host_dy, float* host_dx,
                              const float* host_aux, int64_t n) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(host_x && host_dy && host_dx && host_aux && n >= 1, "bad argument");
    HostStage* s;
    int rc = ensure_stage(ctx, n, true, &s);
    if (rc) return rc;
    B2Q_CHECK_CUDA(cudaMemcpyAsync(s->aux, host_aux, sizeof(float), cudaMemcpyHostToDevice, s->stream));
    B2Q_CHECK_CUDA(cudaMemcpyAsync(s->a, host_x, sizeof(float) * n, cudaMemcpyHostToDevice, s->stream));
    B2Q_CHECK_CUDA(cudaMemcpyAsync(s->b, host_dy, sizeof(float) * n, cudaMemcpyHostToDevice, s->stream));
    rc = b2q_clipgrad_bwd_f32(ctx, s->a, s->b, s->c, s->aux, n, s->stream);
    if (rc) return rc;
    B2Q_CHECK_CUDA(cudaMemcpyAsync(host_dx, s->c, sizeof(float) * n, cudaMemcpyDeviceToHost, s->stream));
    return 0;
}

}  // extern "C"
