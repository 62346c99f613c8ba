// Second-tier operators: CLIP_RELU_PY, WNQ_PY, PACT_PY/PACT_V2_PY, DoReFa_PY,
// QIL_PY/QIL_V2_PY/QIL_V3_PY.  Same numerics contract as the first tier (every reference mx.nd call is one
// separately rounded float32 operation); sums are accumulated in double and rounded once.
#include <cstring>

#include "b2q_common.cuh"
#include "b2q_qdq.cuh"
#include "b2q_reduce.cuh"

#define B2Q_CTX(ctx)                               \
    B2Q_REQUIRE((ctx) != nullptr, "null context"); \
    B2Q_CHECK_CUDA(cudaSetDevice((ctx)->device))

static const Prescale kNoPrescale = {nullptr, nullptr, 0.f};
static const FoldBias kNoBias = {nullptr, nullptr, nullptr};

__device__ __forceinline__ void put(float* dst, int64_t i, float v, int add) {
    dst[i] = add ? __fadd_rn(dst[i], v) : v;
}

__device__ __forceinline__ void put_scalar(float* dst, float v, int req) {
    if (req == B2Q_REQ_NULL || dst == nullptr) return;
    dst[0] = (req == B2Q_REQ_ADD) ? __fadd_rn(dst[0], v) : v;
}

// ------------------------------------------------------------------------------------------------
// Generic flat "elementwise + up to 3 sums + 1 max" kernel.  Op provides
//   __device__ void setup();
//   __device__ EwIn load(int64_t i) const;                      (all global loads of element i)
//   __device__ void apply(int64_t i, const EwIn& in, double* s, float& m);
//   __device__ void finalize(const double* s, float m);         (last block, thread 0)
// Loads of four elements are issued before any of them is consumed (memory-level parallelism).
// ------------------------------------------------------------------------------------------------
struct EwIn {
    float a, b;
};

template <class Op>
__global__ void __launch_bounds__(B2Q_THREADS) ew_kernel(Op op, int64_t n, b2q_slot* slot) {
    b2q_pdl_sync();
    __shared__ double smem[32];
    __shared__ unsigned int s_ticket;
    op.setup();
    double s[3] = {0.0, 0.0, 0.0};
    float m = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x * 4 + threadIdx.x; base < n; base += stride) {
        EwIn in[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t i = base + (int64_t)k * blockDim.x;
            if (i < n) in[k] = op.load(i);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t i = base + (int64_t)k * blockDim.x;
            if (i < n) op.apply(i, in[k], s, m);
        }
    }
    if (!Op::REDUCES) return;
    double r[4];
    r[0] = block_reduce<false>(s[0], smem);
    r[1] = block_reduce<false>(s[1], smem);
    r[2] = block_reduce<false>(s[2], smem);
    r[3] = block_reduce<true>((double)m, smem);
    if (threadIdx.x == 0) {
        for (int k = 0; k < 4; ++k) slot->partial[4 * blockIdx.x + k] = r[k];
        s_ticket = b2q_take_ticket(&slot->ticket, gridDim.x - 1);
    }
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    double a[3] = {0.0, 0.0, 0.0};
    float mm = 0.f;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
        a[0] += __ldcg(&slot->partial[4 * i + 0]);
        a[1] += __ldcg(&slot->partial[4 * i + 1]);
        a[2] += __ldcg(&slot->partial[4 * i + 2]);
        mm = fmaxf(mm, (float)__ldcg(&slot->partial[4 * i + 3]));
    }
    double t[3];
    t[0] = block_reduce<false>(a[0], smem);
    t[1] = block_reduce<false>(a[1], smem);
    t[2] = block_reduce<false>(a[2], smem);
    const float tm = (float)block_reduce<true>((double)mm, smem);
    if (threadIdx.x == 0) {
        op.finalize(t, tm);
        slot->ticket = 0;
    }
}

template <class Op>
static int launch_ew(b2q_ctx* ctx, Op op, int64_t n, cudaStream_t st) {
    B2Q_REQUIRE(n >= 1, "empty tensor");
    int64_t grid = (n + B2Q_THREADS * 4 - 1) / (B2Q_THREADS * 4);
    int64_t cap = Op::REDUCES ? (int64_t)ctx->num_sms * 16 : (int64_t)0x7fffffff;   // reducing ops keep 4 partials per block
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    b2q_launch(ctx, ew_kernel<Op>, (unsigned)grid, B2Q_THREADS, st, op, n, b2q_take_slot(ctx));
    B2Q_LAUNCH_CHECK(ctx);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// WNQ_PY  (core/operator/WNQ.py:51-85)
// ------------------------------------------------------------------------------------------------
struct WnqFwd {
    static const bool REDUCES = false;
    const float* x; float* y; const float* m; int64_t cols; int per_channel; float L; int add;
    __device__ void setup() {}
    __device__ EwIn load(int64_t i) const { return {x[i], 0.f}; }
    __device__ void apply(int64_t i, const EwIn& in, double*, float&) {
        const float mm = m[per_channel ? i / cols : 0];
        const float normed = __fdiv_rn(in.a, mm);                                  // WNQ.py:62
        const float code = roundf(__fmul_rn(normed, L));
        put(y, i, __fmul_rn(__fdiv_rn(code, L), mm), add);                         // :63
    }
    __device__ void finalize(const double*, float) {}
};

// sum_g( dy * x * [|x| != m_g] ) over the (1, groups, cols) view, then max_abs_grad = -sum / m   (:73 / :83)
__global__ void __launch_bounds__(128)
wnq_bwd_sum_kernel(const float* __restrict__ x, const float* __restrict__ dy, SegPlan pl, const float* __restrict__ m,
                   float* __restrict__ mgrad, b2q_slot* slot) {
    __shared__ double smem[32];
    __shared__ unsigned int s_ticket;
    const SegPiece pc = seg_piece(pl);
    const float mm = m[pc.g];
    double acc = 0.0;
    for (int64_t o = pc.o0; o < pc.o1; ++o) {
        const int64_t off = (o * pl.groups + pc.g) * pl.inner;
        for (int64_t i = pc.i0 + threadIdx.x; i < pc.i1; i += blockDim.x) {
            const float xv = x[off + i];
            const float nm = (fabsf(xv) != mm) ? 1.f : 0.f;
            acc += (double)__fmul_rn(__fmul_rn(dy[off + i], xv), nm);
        }
    }
    double r = block_reduce<false>(acc, smem);
    if (threadIdx.x == 0) {
        slot->partial[blockIdx.x] = r;
        s_ticket = b2q_take_ticket(&slot->ticket, gridDim.x - 1);
    }
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    const int SP = pl.S * pl.P;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int64_t gg = wid; gg < pl.groups; gg += nw) {
        double a = 0.0;
        for (int l = lane; l < SP; l += 32) a += __ldcg(&slot->partial[gg * SP + l]);
        a = warp_sum(a);
        if (lane == 0) mgrad[gg] = __fdiv_rn(-(float)a, m[gg]);
    }
    __syncthreads();
    if (threadIdx.x == 0) slot->ticket = 0;
}

struct WnqBwdApply {
    static const bool REDUCES = false;
    const float* x; const float* dy; float* dx; const float* m; const float* mgrad; int64_t cols; int per_channel; int add;
    __device__ void setup() {}
    __device__ EwIn load(int64_t i) const { return {x[i], dy[i]}; }
    __device__ void apply(int64_t i, const EwIn& in, double*, float&) {
        const int64_t g = per_channel ? i / cols : 0;
        const float ax = fabsf(in.a);
        const float nm = (ax != m[g]) ? 1.f : 0.f, im = (ax == m[g]) ? 1.f : 0.f;
        put(dx, i, __fadd_rn(__fmul_rn(in.b, nm), __fmul_rn(mgrad[g], im)), add);  // :85
    }
    __device__ void finalize(const double*, float) {}
};

// ------------------------------------------------------------------------------------------------
// PACT_PY / PACT_V2_PY backward  (core/operator/PACT.py:142-144, 201-203): autograd of mx.nd.where
// ------------------------------------------------------------------------------------------------
struct PactBwd {
    static const bool REDUCES = true;
    const float* x; const float* dy; float* dx; float* dgamma; const float* gamma; int two_sided; int add; int req; int req_gamma;
    float g;
    __device__ void setup() { g = gamma[0]; }
    __device__ EwIn load(int64_t i) const { return {x[i], dy[i]}; }
    __device__ void apply(int64_t i, const EwIn& in, double* s, float&) {
        const float xv = in.a, d = in.b;
        const bool cond = two_sided ? (fabsf(xv) < g) : (xv < g);
        if (req != B2Q_REQ_NULL) put(dx, i, cond ? d : 0.f, add);
        float other = cond ? 0.f : d;
        if (two_sided) other = __fmul_rn(other, mx_sign(xv));
        s[0] += (double)other;
    }
    __device__ void finalize(const double* s, float) { put_scalar(dgamma, (float)s[0], req_gamma); }
};

// ------------------------------------------------------------------------------------------------
// DoReFa_PY  (core/operator/PACT.py:44-50, 76-77)
// ------------------------------------------------------------------------------------------------
struct DorefaMax {
    static const bool REDUCES = true;
    const float* x; float* vmax;
    __device__ void setup() {}
    __device__ EwIn load(int64_t i) const { return {x[i], 0.f}; }
    __device__ void apply(int64_t i, const EwIn& in, double*, float& m) { m = fmaxf(m, fabsf(tanhf(in.a))); }
    __device__ void finalize(const double*, float m) { vmax[0] = m; }
};

struct DorefaFwd {
    static const bool REDUCES = false;
    const float* x; float* y; const float* vmax; float L; int add;
    float two_v;
    __device__ void setup() { two_v = __fmul_rn(2.f, vmax[0]); }
    __device__ EwIn load(int64_t i) const { return {x[i], 0.f}; }
    __device__ void apply(int64_t i, const EwIn& in, double*, float&) {
        const float t = tanhf(in.a);
        const float o = __fadd_rn(__fdiv_rn(t, two_v), 0.5f);                      // PACT.py:49
        const float code = roundf(__fmul_rn(L, o));                               // quantizeK, :26-28
        put(y, i, __fsub_rn(__fmul_rn(2.f, __fdiv_rn(code, L)), 1.f), add);        // :50
    }
    __device__ void finalize(const double*, float) {}
};

struct DorefaBwdSum {  // d(2v) = sum( g * (-t / (2v)^2) ),  g = 2*dy
    static const bool REDUCES = true;
    const float* x; const float* dy; const float* vmax; float* dv_out;
    float two_v, sq;
    __device__ void setup() { two_v = __fmul_rn(2.f, vmax[0]); sq = __fmul_rn(two_v, two_v); }
    __device__ EwIn load(int64_t i) const { return {x[i], dy[i]}; }
    __device__ void apply(int64_t i, const EwIn& in, double* s, float&) {
        const float t = tanhf(in.a);
        const float g = __fmul_rn(2.f, in.b);
        s[0] += (double)__fmul_rn(g, __fdiv_rn(-t, sq));
    }
    __device__ void finalize(const double* s, float) { dv_out[0] = __fmul_rn(2.f, (float)s[0]); }
};

struct DorefaBwdApply {
    static const bool REDUCES = false;
    const float* x; const float* dy; float* dx; const float* vmax; const float* dv; int add;
    float v, two_v, dvv;
    __device__ void setup() { v = vmax[0]; two_v = __fmul_rn(2.f, v); dvv = dv[0]; }
    __device__ EwIn load(int64_t i) const { return {x[i], dy[i]}; }
    __device__ void apply(int64_t i, const EwIn& in, double*, float&) {
        const float t = tanhf(in.a);
        const float g = __fmul_rn(2.f, in.b);
        float dt = __fdiv_rn(g, two_v);
        const float ismax = (fabsf(t) == v) ? 1.f : 0.f;
        dt = __fadd_rn(dt, __fmul_rn(__fmul_rn(ismax, dvv), mx_sign(t)));
        put(dx, i, __fmul_rn(dt, __fsub_rn(1.f, __fmul_rn(t, t))), add);
    }
    __device__ void finalize(const double*, float) {}
};

// ------------------------------------------------------------------------------------------------
// QIL_PY / QIL_V2_PY / QIL_V3_PY  (core/operator/QIL.py, QIL_V2.py, QIL_V3.py)
// ------------------------------------------------------------------------------------------------
struct QilParams {
    float pp, cp, a, b, center, distance;
};

__device__ __forceinline__ QilParams qil_params(int variant, float p0, float p1) {
    QilParams q;
    if (variant == 1) {           // QIL.py:72-75
        q.pp = p0; q.cp = p1;
        q.center = __fmul_rn(0.5f, __fadd_rn(p1, p0));
        q.distance = __fmul_rn(0.5f, __fsub_rn(p1, p0));
    } else if (variant == 2) {    // QIL_V2.py:46-49
        q.center = p0; q.distance = p1;
        q.cp = __fadd_rn(p0, p1); q.pp = __fsub_rn(p0, p1);
    } else {                      // QIL_V3.py:50-52
        q.pp = expf(p0); q.distance = expf(p1);
        q.cp = __fadd_rn(q.pp, q.distance);
        q.center = 0.f;
    }
    q.a = __fdiv_rn(0.5f, q.distance);
    q.b = __fadd_rn(__fdiv_rn(__fmul_rn(-0.5f, q.center), q.distance), 0.5f);
    return q;
}

__global__ void qil_v1_clamp_kernel(float* p0, float* p1) {   // QIL.py:51-54
    if (p0[0] < 0.f) p0[0] = 0.f;
    if (p1[0] > 1.f) p1[0] = 1.f;
}

struct QilFwd {
    static const bool REDUCES = false;
    int variant; const float* x; float* y; const float* p0; const float* p1; float L; int add;
    QilParams q;
    __device__ void setup() { q = qil_params(variant, p0[0], p1[0]); }
    __device__ EwIn load(int64_t i) const { return {x[i], 0.f}; }
    __device__ void apply(int64_t i, const EwIn& in, double*, float&) {
        const float xv = in.a, ax = fabsf(xv), sg = mx_sign(xv);
        const float inter = __fmul_rn((ax >= q.pp) ? 1.f : 0.f, (ax <= q.cp) ? 1.f : 0.f);
        const float lin = (variant == 3) ? __fdiv_rn(__fsub_rn(ax, q.pp), q.distance)
                                         : __fadd_rn(__fmul_rn(q.a, ax), q.b);
        const float out = __fadd_rn(__fmul_rn(sg, (ax > q.cp) ? 1.f : 0.f), __fmul_rn(__fmul_rn(sg, lin), inter));
        const float code = roundf(__fmul_rn(out, L));
        put(y, i, __fdiv_rn(code, L), add);
    }
    __device__ void finalize(const double*, float) {}
};

struct QilBwd {
    static const bool REDUCES = true;
    int variant; const float* x; const float* dy; float* dx; const float* p0; const float* p1; float* dp0; float* dp1;
    int req, add, req_p0, req_p1;
    QilParams q;
    __device__ void setup() { q = qil_params(variant, p0[0], p1[0]); }
    __device__ EwIn load(int64_t i) const { return {x[i], dy[i]}; }
    __device__ void apply(int64_t i, const EwIn& in, double* s, float&) {
        const float xv = in.a, ax = fabsf(xv), sg = mx_sign(xv);
        const float inter = __fmul_rn((ax >= q.pp) ? 1.f : 0.f, (ax <= q.cp) ? 1.f : 0.f);
        const float g = __fmul_rn(__fmul_rn(in.b, sg), inter);     // d out / d lin
        float d;
        if (variant == 3) {
            d = __fmul_rn(__fdiv_rn(g, q.distance), sg);
            s[0] += (double)__fdiv_rn(-g, q.distance);
            s[1] += (double)__fmul_rn(g, __fdiv_rn(-__fsub_rn(ax, q.pp), __fmul_rn(q.distance, q.distance)));
        } else {
            d = __fmul_rn(__fmul_rn(g, q.a), sg);
            s[0] += (double)__fmul_rn(g, ax);   // d/da
            s[1] += (double)g;                  // d/db
        }
        if (req != B2Q_REQ_NULL) put(dx, i, d, add);
    }
    __device__ void finalize(const double* s, float) {
        if (variant == 3) {
            put_scalar(dp0, __fmul_rn((float)s[0], q.pp), req_p0);
            put_scalar(dp1, __fmul_rn((float)s[1], q.distance), req_p1);
            return;
        }
        const float da = (float)s[0], db = (float)s[1];
        const float dsq = __fmul_rn(q.distance, q.distance);
        const float dd = __fadd_rn(__fmul_rn(da, __fdiv_rn(-0.5f, dsq)),
                                   __fmul_rn(db, __fdiv_rn(__fmul_rn(0.5f, q.center), dsq)));
        const float dc = __fmul_rn(db, __fdiv_rn(-0.5f, q.distance));
        if (variant == 2) {
            put_scalar(dp0, dc, req_p0);
            put_scalar(dp1, dd, req_p1);
        } else {
            put_scalar(dp1, __fadd_rn(__fmul_rn(0.5f, dc), __fmul_rn(0.5f, dd)), req_p1);   // clipping point
            put_scalar(dp0, __fsub_rn(__fmul_rn(0.5f, dc), __fmul_rn(0.5f, dd)), req_p0);   // pruning point
        }
    }
};

// ------------------------------------------------------------------------------------------------
// True-int8 export: the integer codes the QDQ sweep computes, packed as int8, plus
// the per-group step q = T / qlevel, so that an inference engine can run the convolution on int8 / fp8 tensor cores.
// codes[i] = clamp(roundf(clip(x[i]) / q), -128, 127); dequantised value = codes[i] * q (bit-identical to the
// fake-quant output whenever |code| <= 127, which holds for every clipping operator and for weights).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(B2Q_THREADS)
export_int8_kernel(const float* __restrict__ x, int8_t* __restrict__ codes, float* __restrict__ steps, int64_t outer,
                   int64_t groups, int64_t inner, const float* __restrict__ thr, float qlevel, int clip_mode, int fast) {
    const int64_t n = outer * groups * inner;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t g = (groups == 1) ? 0 : (i / inner) % groups;
        const float T = thr[g];
        const QScale s = make_qscale(T, qlevel, fast != 0);
        const float c = quant_code(clip_value(clip_mode, x[i], T), s);
        codes[i] = (int8_t)max(-128, min(127, __float2int_rn(c)));
        if (i < groups && steps) steps[i] = make_qscale(thr[i], qlevel, false).q;
    }
}

extern "C" {

int b2q_clip_relu_fwd_f32(b2q_ctx* ctx, const float* x, float* y, int64_t n, float threshold, float q, int req,
                          void* stream) {
    B2Q_CTX(ctx);
    if (req == B2Q_REQ_NULL) return 0;
    B2Q_REQUIRE(x && y && n >= 1, "bad argument");
    // GDRQ.py:202-204: clip(x, 0, thr); q = thr/L computed in python double, applied in float32
    QdqArgs a = {nullptr, nullptr, q, threshold, 0.f, ctx->fast_div, nullptr, B2Q_CLIP_ZERO_T, 1, req};
    return launch_qdq(ctx, x, y, 1, 1, n, kNoPrescale, kNoBias, a, (cudaStream_t)stream);
}

int b2q_export_int8_f32(b2q_ctx* ctx, const float* x, int8_t* codes, float* steps, int64_t outer, int64_t groups,
                        int64_t inner, const float* thr, float qlevel, int clip_mode, void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(x && codes && thr && outer >= 1 && groups >= 1 && inner >= 1, "bad argument");
    B2Q_REQUIRE(clip_mode >= B2Q_CLIP_NONE && clip_mode <= B2Q_CLIP_WHERE_LT, "unknown clip_mode");
    const int64_t n = outer * groups * inner;
    int64_t grid = (n + B2Q_THREADS - 1) / B2Q_THREADS;
    if (grid > (int64_t)ctx->num_sms * 32) grid = (int64_t)ctx->num_sms * 32;
    b2q_timed_launch tl(ctx, B2Q_KIND_OTHER, 5.0 * (double)n, (cudaStream_t)stream);
    export_int8_kernel<<<(unsigned)grid, B2Q_THREADS, 0, (cudaStream_t)stream>>>(x, codes, steps, outer, groups, inner, thr,
                                                                                 qlevel, clip_mode, ctx->fast_div);
    B2Q_LAUNCH_CHECK(ctx);
    return 0;
}

int b2q_wnq_fwd_f32(b2q_ctx* ctx, const float* x, float* y, int64_t rows, int64_t cols, int per_channel, float qlevel,
                    int req, void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(x && y && rows >= 1 && cols >= 1, "bad argument");
    if (req == B2Q_REQ_NULL) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    b2q_slot* slot = b2q_take_slot(ctx);
    UpdateArgs u;
    memset(&u, 0, sizeof(u));
    u.stat_out = slot->scale;
    const int64_t groups = per_channel ? rows : 1, inner = per_channel ? cols : rows * cols;
    int rc = launch_reduce<true>(ctx, slot, x, 1, groups, inner, kNoPrescale, u, st);
    if (rc) return rc;
    WnqFwd op = {x, y, slot->scale, cols, per_channel, qlevel, req == B2Q_REQ_ADD};
    return launch_ew(ctx, op, rows * cols, st);
}

int b2q_wnq_bwd_f32(b2q_ctx* ctx, const float* x, const float* dy, float* dx, int64_t rows, int64_t cols,
                    int per_channel, int req, void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(x && dy && dx && rows >= 1 && cols >= 1, "bad argument");
    if (req == B2Q_REQ_NULL) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    b2q_slot* slot = b2q_take_slot(ctx);
    UpdateArgs u;
    memset(&u, 0, sizeof(u));
    u.stat_out = slot->scale;
    const int64_t groups = per_channel ? rows : 1, inner = per_channel ? cols : rows * cols;
    int rc = launch_reduce<true>(ctx, slot, x, 1, groups, inner, kNoPrescale, u, st);
    if (rc) return rc;
    SegPlan pl = b2q_seg_plan(nullptr, nullptr, 1, groups, inner, ctx->num_sms * 8);
    wnq_bwd_sum_kernel<<<(unsigned)(groups * pl.S * pl.P), 128, 0, st>>>(x, dy, pl, slot->scale, slot->clip, slot);
    B2Q_LAUNCH_CHECK(ctx);
    WnqBwdApply op = {x, dy, dx, slot->scale, slot->clip, cols, per_channel, req == B2Q_REQ_ADD};
    return launch_ew(ctx, op, rows * cols, st);
}

int b2q_pact_bwd_f32(b2q_ctx* ctx, const float* x, const float* dy, float* dx, float* dgamma, const float* gamma,
                     int64_t n, int two_sided, int req, int req_gamma, void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(x && dy && gamma && n >= 1, "bad argument");
    B2Q_REQUIRE(req == B2Q_REQ_NULL || dx, "null dx");
    PactBwd op = {x, dy, dx, dgamma, gamma, two_sided, req == B2Q_REQ_ADD, req, req_gamma, 0.f};
    return launch_ew(ctx, op, n, (cudaStream_t)stream);
}

int b2q_dorefa_fwd_f32(b2q_ctx* ctx, const float* x, float* y, float* vmax_out, int64_t n, float qlevel, int req,
                       void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(x && y && vmax_out && n >= 1, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    DorefaMax mop = {x, vmax_out};
    int rc = launch_ew(ctx, mop, n, st);
    if (rc || req == B2Q_REQ_NULL) return rc;
    DorefaFwd op = {x, y, vmax_out, qlevel, req == B2Q_REQ_ADD, 0.f};
    return launch_ew(ctx, op, n, st);
}

int b2q_dorefa_bwd_f32(b2q_ctx* ctx, const float* x, const float* dy, float* dx, const float* vmax, int64_t n, int req,
                       void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(x && dy && dx && vmax && n >= 1, "bad argument");
    if (req == B2Q_REQ_NULL) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    b2q_slot* slot = b2q_take_slot(ctx);
    DorefaBwdSum sop = {x, dy, vmax, slot->scale, 0.f, 0.f};
    int rc = launch_ew(ctx, sop, n, st);
    if (rc) return rc;
    DorefaBwdApply op = {x, dy, dx, vmax, slot->scale, req == B2Q_REQ_ADD, 0.f, 0.f, 0.f};
    return launch_ew(ctx, op, n, st);
}

int b2q_qil_fwd_f32(b2q_ctx* ctx, int variant, const float* x, float* y, float* p0, float* p1, int64_t n, float qlevel,
                    int req, void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(variant >= 1 && variant <= 3, "variant must be 1, 2 or 3");
    B2Q_REQUIRE(x && y && p0 && p1 && n >= 1, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (variant == 1) {
        qil_v1_clamp_kernel<<<1, 1, 0, st>>>(p0, p1);
        B2Q_LAUNCH_CHECK(ctx);
    }
    if (req == B2Q_REQ_NULL) return 0;
    QilFwd op = {variant, x, y, p0, p1, qlevel, req == B2Q_REQ_ADD, {}};
    return launch_ew(ctx, op, n, st);
}

int b2q_qil_bwd_f32(b2q_ctx* ctx, int variant, const float* x, const float* dy, float* dx, const float* p0,
                    const float* p1, float* dp0, float* dp1, int64_t n, int req, int req_p0, int req_p1, void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(variant >= 1 && variant <= 3, "variant must be 1, 2 or 3");
    B2Q_REQUIRE(x && dy && p0 && p1 && n >= 1, "bad argument");
    QilBwd op = {variant, x, dy, dx, p0, p1, dp0, dp1, req, req == B2Q_REQ_ADD, req_p0, req_p1, {}};
    return launch_ew(ctx, op, n, (cudaStream_t)stream);
}

}  // extern "C"
