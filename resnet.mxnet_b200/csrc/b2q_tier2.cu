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

__device__ __forceinline__ void put_scalar(float* dst, float v, int req) {
    if (req == B2Q_REQ_NULL || dst == nullptr) return;
    dst[0] = (req == B2Q_REQ_ADD) ? __fadd_rn(dst[0], v) : v;
}

// c / L for an integer-valued code c and a level count L = 2^nbits - 1, bit-identical to IEEE division
// (mx.nd.round(x * L) / L: QIL.py:99, PACT.py:26-28, WNQ.py:63) in three instructions instead of the ~10 of the
// division subroutine: q0 = RN(c * rL), r = c - q0 * L (exact in one FMA), q1 = RN(q0 + r * rL) with rL = RN(1/L) is
// the correctly rounded quotient (Markstein); the sign of a zero quotient is taken from c.  NaN / Inf / |c| > 2^24 take
// the reference division.  b2q_selftest(1) compares it with __fdiv_rn for every integer |c| <= 4 L, nbits = 1..16.
struct LevelDiv {
    float L, rL;
};

__device__ __forceinline__ LevelDiv make_level_div(float L) { return {L, __frcp_rn(L)}; }

__device__ __forceinline__ float div_level(float c, const LevelDiv& d) {
    const float q0 = __fmul_rn(c, d.rL);
    const float r = __fmaf_rn(-q0, d.L, c);
    const float q1 = copysignf(__fmaf_rn(r, d.rL, q0), c);
    return (fabsf(c) <= 16777216.f) ? q1 : __fdiv_rn(c, d.L);
}

// ------------------------------------------------------------------------------------------------
// Generic flat "elementwise + up to 2 sums + 2 maxima" kernel over 256-bit words.  Op provides
//   static const int NIN (1: x, 2: x and dy);  static const bool REDUCES;
//   const float* in0() / in1();  float* out() (nullptr: nothing is written);  int add() (req == add)
//   __device__ void setup();                                     (per-thread scalars from device memory)
//   __device__ float apply(float a, float b, int64_t i, double* s, float* m);   (returns the output value; s[2], m[2])
//   __device__ void finalize(const double* s, const float* m);   (last block, thread 0)
// Same structure as the hot sweeps: head scalars until 32-byte alignment, two 256-bit loads per input in flight per
// thread, one 256-bit store per word, tail scalars; reducing ops run a capped grid-stride grid and finish in the block
// that draws the last ticket.
// ------------------------------------------------------------------------------------------------
#define B2Q_EW_UNROLL 2

template <class Op>
__device__ __forceinline__ void ew_scalar(Op& op, int64_t i, double* s, float* m) {
    const float a = op.in0()[i];
    const float b = (Op::NIN == 2) ? op.in1()[i] : 0.f;
    const float r = op.apply(a, b, i, s, m);
    float* o = op.out();
    if (o) o[i] = op.add() ? __fadd_rn(o[i], r) : r;
}

template <class Op, bool VEC>
__global__ void __launch_bounds__(B2Q_THREADS) ew_kernel(Op op, FlatSplit sp, int64_t n, b2q_slot* slot) {
    b2q_pdl_sync();
    __shared__ double smem[32];
    __shared__ unsigned int s_ticket;
    op.setup();
    double s[2] = {0.0, 0.0};
    float m[2] = {0.f, 0.f};
    if (VEC) {
        const float* xa = op.in0() + sp.head;
        const float* xb = (Op::NIN == 2) ? op.in1() + sp.head : nullptr;
        float* yo = op.out() ? op.out() + sp.head : nullptr;
        const bool add = op.add() != 0;
        const int64_t tile = (int64_t)B2Q_THREADS * B2Q_EW_UNROLL;
        const int64_t ntiles = (sp.n8 + tile - 1) / tile;
        for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int64_t base = t * tile + threadIdx.x;
            f8 va[B2Q_EW_UNROLL], vb[B2Q_EW_UNROLL], vo[B2Q_EW_UNROLL];
#pragma unroll
            for (int k = 0; k < B2Q_EW_UNROLL; ++k) {
                const int64_t w = base + (int64_t)k * B2Q_THREADS;
                if (w < sp.n8) {
                    va[k] = ld_f8<2>(xa + 8 * w);
                    if (Op::NIN == 2) vb[k] = ld_f8<2>(xb + 8 * w);
                    if (add && yo) vo[k] = ld_f8<0>(yo + 8 * w);
                }
            }
#pragma unroll
            for (int k = 0; k < B2Q_EW_UNROLL; ++k) {
                const int64_t w = base + (int64_t)k * B2Q_THREADS;
                if (w < sp.n8) {
                    f8 r;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        r.v[j] = op.apply(va[k].v[j], (Op::NIN == 2) ? vb[k].v[j] : 0.f, sp.head + 8 * w + j, s, m);
                        if (add && yo) r.v[j] = __fadd_rn(vo[k].v[j], r.v[j]);
                    }
                    if (yo) st_f8<0>(yo + 8 * w, r);
                }
            }
        }
        if (blockIdx.x == 0) {   // the (at most 14) unaligned scalars
            const int64_t tid = threadIdx.x;
            if (tid < sp.head) ew_scalar(op, tid, s, m);
            else if (tid - sp.head < sp.tail) ew_scalar(op, sp.head + 8 * sp.n8 + (tid - sp.head), s, m);
        }
    } else {   // mutually misaligned buffers: scalar accesses, four elements in flight per thread
        const int64_t stride = (int64_t)gridDim.x * blockDim.x;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) ew_scalar(op, i, s, m);
    }
    if (!Op::REDUCES) return;
    double r[4];
    r[0] = block_reduce<false>(s[0], smem);
    r[1] = block_reduce<false>(s[1], smem);
    r[2] = block_reduce<true>((double)m[0], smem);
    r[3] = block_reduce<true>((double)m[1], smem);
    if (gridDim.x == 1) {
        if (threadIdx.x == 0) {
            const float fm[2] = {(float)r[2], (float)r[3]};
            op.finalize(r, fm);
        }
        return;
    }
    if (threadIdx.x == 0) {
        for (int k = 0; k < 4; ++k) slot->partial[4 * blockIdx.x + k] = r[k];
        s_ticket = b2q_take_ticket(&slot->ticket, gridDim.x - 1);
    }
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    double a[2] = {0.0, 0.0};
    float mm[2] = {0.f, 0.f};
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
        a[0] += __ldcg(&slot->partial[4 * i + 0]);
        a[1] += __ldcg(&slot->partial[4 * i + 1]);
        mm[0] = fmax_nan(mm[0], (float)__ldcg(&slot->partial[4 * i + 2]));
        mm[1] = fmax_nan(mm[1], (float)__ldcg(&slot->partial[4 * i + 3]));
    }
    double t[2];
    t[0] = block_reduce<false>(a[0], smem);
    t[1] = block_reduce<false>(a[1], smem);
    float tm[2];
    tm[0] = (float)block_reduce<true>((double)mm[0], smem);
    tm[1] = (float)block_reduce<true>((double)mm[1], smem);
    if (threadIdx.x == 0) {
        op.finalize(t, tm);
        slot->ticket = 0;
    }
}

template <class Op>
static int launch_ew(b2q_ctx* ctx, Op op, int64_t n, double alg_bytes_per_elem, cudaStream_t st) {
    B2Q_REQUIRE(n >= 1, "empty tensor");
    FlatSplit sp = b2q_flat_split(op.in0(), n);
    bool vec = sp.head <= B2Q_THREADS;
    if (Op::NIN == 2) vec = vec && same_misalignment(op.in0(), op.in1());
    if (op.out()) vec = vec && same_misalignment(op.in0(), op.out());
    const int64_t cap = Op::REDUCES ? (int64_t)ctx->num_sms * 16 : (int64_t)0x7fffffff;   // 4 partials per block
    b2q_timed_launch tl(ctx, B2Q_KIND_OTHER, alg_bytes_per_elem * (double)n, st);
    if (vec) {
        const int64_t tile = (int64_t)B2Q_THREADS * B2Q_EW_UNROLL;
        int64_t grid = (sp.n8 + tile - 1) / tile;
        if (grid > cap) grid = cap;
        if (grid < 1) grid = 1;
        b2q_launch(ctx, ew_kernel<Op, true>, (unsigned)grid, B2Q_THREADS, st, op, sp, n, b2q_take_slot(ctx, st));
    } else {
        int64_t grid = (n + B2Q_THREADS * 4 - 1) / (B2Q_THREADS * 4);
        if (grid > cap) grid = cap;
        if (grid < 1) grid = 1;
        b2q_launch(ctx, ew_kernel<Op, false>, (unsigned)grid, B2Q_THREADS, st, op, sp, n, b2q_take_slot(ctx, st));
    }
    B2Q_LAUNCH_CHECK(ctx);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// WNQ_PY  (core/operator/WNQ.py:51-85)
// ------------------------------------------------------------------------------------------------
template <bool PER_CHANNEL>
struct WnqFwd {
    static const int NIN = 1;
    static const bool REDUCES = false;
    const float* x; float* y; const float* m; int64_t cols; float L; int add_;
    float m0; LevelDiv ld;
    __host__ __device__ const float* in0() const { return x; }
    __host__ __device__ const float* in1() const { return nullptr; }
    __host__ __device__ float* out() const { return y; }
    __host__ __device__ int add() const { return add_; }
    __device__ void setup() { m0 = m[0]; ld = make_level_div(L); }
    __device__ float apply(float a, float, int64_t i, double*, float*) {
        const float mm = PER_CHANNEL ? m[i / cols] : m0;
        const float normed = __fdiv_rn(a, mm);                                     // WNQ.py:62
        const float code = roundf(__fmul_rn(normed, L));
        return __fmul_rn(div_level(code, ld), mm);                                 // :63
    }
    __device__ void finalize(const double*, const float*) {}
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
        if (pl.vec == 4) {   // 128-bit loads, two of each tensor in flight per thread
            const float4* x4 = reinterpret_cast<const float4*>(x + off);
            const float4* g4 = reinterpret_cast<const float4*>(dy + off);
            const int64_t end = pc.i1 >> 2;
            for (int64_t i = (pc.i0 >> 2) + threadIdx.x; i < end; i += 2 * (int64_t)blockDim.x) {
                const int64_t i2 = i + blockDim.x;
                const float4 xa = x4[i], ga = g4[i];
                float4 xb = make_float4(0.f, 0.f, 0.f, 0.f), gb = xb;
                if (i2 < end) { xb = x4[i2]; gb = g4[i2]; }
#define B2Q_WNQ_TERM(X, G) (double)__fmul_rn(__fmul_rn(G, X), (fabsf(X) != mm) ? 1.f : 0.f)
                acc += (B2Q_WNQ_TERM(xa.x, ga.x) + B2Q_WNQ_TERM(xa.y, ga.y)) + (B2Q_WNQ_TERM(xa.z, ga.z) + B2Q_WNQ_TERM(xa.w, ga.w));
                if (i2 < end)
                    acc += (B2Q_WNQ_TERM(xb.x, gb.x) + B2Q_WNQ_TERM(xb.y, gb.y)) + (B2Q_WNQ_TERM(xb.z, gb.z) + B2Q_WNQ_TERM(xb.w, gb.w));
#undef B2Q_WNQ_TERM
            }
            continue;
        }
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

template <bool PER_CHANNEL>
struct WnqBwdApply {
    static const int NIN = 2;
    static const bool REDUCES = false;
    const float* x; const float* dy; float* dx; const float* m; const float* mgrad; int64_t cols; int add_;
    float m0, g0;
    __host__ __device__ const float* in0() const { return x; }
    __host__ __device__ const float* in1() const { return dy; }
    __host__ __device__ float* out() const { return dx; }
    __host__ __device__ int add() const { return add_; }
    __device__ void setup() { m0 = m[0]; g0 = mgrad[0]; }
    __device__ float apply(float a, float b, int64_t i, double*, float*) {
        const int64_t g = PER_CHANNEL ? i / cols : 0;
        const float mm = PER_CHANNEL ? m[g] : m0, mg = PER_CHANNEL ? mgrad[g] : g0;
        const float ax = fabsf(a);
        const float nm = (ax != mm) ? 1.f : 0.f, im = (ax == mm) ? 1.f : 0.f;
        return __fadd_rn(__fmul_rn(b, nm), __fmul_rn(mg, im));                     // :85
    }
    __device__ void finalize(const double*, const float*) {}
};

// ------------------------------------------------------------------------------------------------
// PACT_PY / PACT_V2_PY backward  (core/operator/PACT.py:142-144, 201-203): autograd of mx.nd.where
// ------------------------------------------------------------------------------------------------
struct PactBwd {
    static const int NIN = 2;
    static const bool REDUCES = true;
    const float* x; const float* dy; float* dx; float* dgamma; const float* gamma; int two_sided; int add_; int req_gamma;
    float g;
    __host__ __device__ const float* in0() const { return x; }
    __host__ __device__ const float* in1() const { return dy; }
    __host__ __device__ float* out() const { return dx; }
    __host__ __device__ int add() const { return add_; }
    __device__ void setup() { g = gamma[0]; }
    __device__ float apply(float xv, float d, int64_t, double* s, float*) {
        const bool cond = two_sided ? (fabsf(xv) < g) : (xv < g);
        float other = cond ? 0.f : d;
        if (two_sided) other = __fmul_rn(other, mx_sign(xv));
        s[0] += (double)other;
        return cond ? d : 0.f;
    }
    __device__ void finalize(const double* s, const float*) { put_scalar(dgamma, (float)s[0], req_gamma); }
};

// ------------------------------------------------------------------------------------------------
// DoReFa_PY  (core/operator/PACT.py:44-50, 76-77)
// max|tanh(w)| without a tanh per element: tanhf is odd and monotonic non-decreasing over the positive floats EXCEPT at
// the switch between its two branches -- on this toolchain exactly one adjacent pair, tanhf(0x3f199999) >
// tanhf(0x3f19999a) (|x| = 0.6, b2q_selftest(3)).  So the first pass takes max|w| (m0) and, for the few elements whose
// |w| lies in a 129-ulp window W around that point, max tanhf(|w|) (m1, evaluated only there); max|tanh(w)| =
// max(tanhf(m0), m1) exactly: outside W tanhf is monotonic, and every value in W is below tanhf of the first float
// above W (b2q_selftest(2) checks both facts over all 2^31 positive floats).  Option "dorefa_tanh_max"=1 restores the
// element-wise max of tanh.
// ------------------------------------------------------------------------------------------------
#define B2Q_TANH_WINDOW_LO 0x3f199959u
#define B2Q_TANH_WINDOW_HI 0x3f1999d9u

struct DorefaAbsMax {
    static const int NIN = 1;
    static const bool REDUCES = true;
    const float* x; float* vmax;
    __host__ __device__ const float* in0() const { return x; }
    __host__ __device__ const float* in1() const { return nullptr; }
    __host__ __device__ float* out() const { return nullptr; }
    __host__ __device__ int add() const { return 0; }
    __device__ void setup() {}
    __device__ float apply(float a, float, int64_t, double*, float* m) {
        const float ax = fabsf(a);
        m[0] = fmax_nan(m[0], ax);
        const unsigned int b = __float_as_uint(ax);
        if (b >= B2Q_TANH_WINDOW_LO && b <= B2Q_TANH_WINDOW_HI) m[1] = fmax_nan(m[1], tanhf(ax));
        return 0.f;
    }
    __device__ void finalize(const double*, const float* m) { vmax[0] = fmax_nan(fabsf(tanhf(m[0])), m[1]); }
};

struct DorefaMax {
    static const int NIN = 1;
    static const bool REDUCES = true;
    const float* x; float* vmax;
    __host__ __device__ const float* in0() const { return x; }
    __host__ __device__ const float* in1() const { return nullptr; }
    __host__ __device__ float* out() const { return nullptr; }
    __host__ __device__ int add() const { return 0; }
    __device__ void setup() {}
    __device__ float apply(float a, float, int64_t, double*, float* m) { m[0] = fmax_nan(m[0], fabsf(tanhf(a))); return 0.f; }
    __device__ void finalize(const double*, const float* m) { vmax[0] = m[0]; }
};

struct DorefaFwd {
    static const int NIN = 1;
    static const bool REDUCES = false;
    const float* x; float* y; const float* vmax; float L; int add_;
    float two_v; LevelDiv ld;
    __host__ __device__ const float* in0() const { return x; }
    __host__ __device__ const float* in1() const { return nullptr; }
    __host__ __device__ float* out() const { return y; }
    __host__ __device__ int add() const { return add_; }
    __device__ void setup() { two_v = __fmul_rn(2.f, vmax[0]); ld = make_level_div(L); }
    __device__ float apply(float a, float, int64_t, double*, float*) {
        const float t = tanhf(a);
        const float o = __fadd_rn(__fdiv_rn(t, two_v), 0.5f);                      // PACT.py:49
        const float code = roundf(__fmul_rn(L, o));                               // quantizeK, :26-28
        return __fsub_rn(__fmul_rn(2.f, div_level(code, ld)), 1.f);                // :50
    }
    __device__ void finalize(const double*, const float*) {}
};

struct DorefaBwdSum {  // d(2v) = sum( g * (-t / (2v)^2) ),  g = 2*dy
    static const int NIN = 2;
    static const bool REDUCES = true;
    const float* x; const float* dy; const float* vmax; float* dv_out;
    float two_v, sq;
    __host__ __device__ const float* in0() const { return x; }
    __host__ __device__ const float* in1() const { return dy; }
    __host__ __device__ float* out() const { return nullptr; }
    __host__ __device__ int add() const { return 0; }
    __device__ void setup() { two_v = __fmul_rn(2.f, vmax[0]); sq = __fmul_rn(two_v, two_v); }
    __device__ float apply(float a, float b, int64_t, double* s, float*) {
        const float t = tanhf(a);
        const float g = __fmul_rn(2.f, b);
        s[0] += (double)__fmul_rn(g, __fdiv_rn(-t, sq));
        return 0.f;
    }
    __device__ void finalize(const double* s, const float*) { dv_out[0] = __fmul_rn(2.f, (float)s[0]); }
};

struct DorefaBwdApply {
    static const int NIN = 2;
    static const bool REDUCES = false;
    const float* x; const float* dy; float* dx; const float* vmax; const float* dv; int add_;
    float v, two_v, dvv;
    __host__ __device__ const float* in0() const { return x; }
    __host__ __device__ const float* in1() const { return dy; }
    __host__ __device__ float* out() const { return dx; }
    __host__ __device__ int add() const { return add_; }
    __device__ void setup() { v = vmax[0]; two_v = __fmul_rn(2.f, v); dvv = dv[0]; }
    __device__ float apply(float a, float b, int64_t, double*, float*) {
        const float t = tanhf(a);
        const float g = __fmul_rn(2.f, b);
        float dt = __fdiv_rn(g, two_v);
        const float ismax = (fabsf(t) == v) ? 1.f : 0.f;
        dt = __fadd_rn(dt, __fmul_rn(__fmul_rn(ismax, dvv), mx_sign(t)));
        return __fmul_rn(dt, __fsub_rn(1.f, __fmul_rn(t, t)));
    }
    __device__ void finalize(const double*, const float*) {}
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

template <int VARIANT3>
struct QilFwd {
    static const int NIN = 1;
    static const bool REDUCES = false;
    int variant; const float* x; float* y; const float* p0; const float* p1; float L; int add_;
    QilParams q; LevelDiv ld;
    __host__ __device__ const float* in0() const { return x; }
    __host__ __device__ const float* in1() const { return nullptr; }
    __host__ __device__ float* out() const { return y; }
    __host__ __device__ int add() const { return add_; }
    __device__ void setup() { q = qil_params(variant, p0[0], p1[0]); ld = make_level_div(L); }
    __device__ float apply(float xv, float, int64_t, double*, float*) {
        const float ax = fabsf(xv), sg = mx_sign(xv);
        const float inter = __fmul_rn((ax >= q.pp) ? 1.f : 0.f, (ax <= q.cp) ? 1.f : 0.f);
        const float lin = VARIANT3 ? __fdiv_rn(__fsub_rn(ax, q.pp), q.distance)
                                   : __fadd_rn(__fmul_rn(q.a, ax), q.b);
        const float out = __fadd_rn(__fmul_rn(sg, (ax > q.cp) ? 1.f : 0.f), __fmul_rn(__fmul_rn(sg, lin), inter));
        const float code = roundf(__fmul_rn(out, L));
        return div_level(code, ld);
    }
    __device__ void finalize(const double*, const float*) {}
};

struct QilBwd {
    static const int NIN = 2;
    static const bool REDUCES = true;
    int variant; const float* x; const float* dy; float* dx; const float* p0; const float* p1; float* dp0; float* dp1;
    int add_, req_p0, req_p1;
    QilParams q; float dsq;
    __host__ __device__ const float* in0() const { return x; }
    __host__ __device__ const float* in1() const { return dy; }
    __host__ __device__ float* out() const { return dx; }
    __host__ __device__ int add() const { return add_; }
    __device__ void setup() { q = qil_params(variant, p0[0], p1[0]); dsq = __fmul_rn(q.distance, q.distance); }
    __device__ float apply(float xv, float gy, int64_t, double* s, float*) {
        const float ax = fabsf(xv), sg = mx_sign(xv);
        const float inter = __fmul_rn((ax >= q.pp) ? 1.f : 0.f, (ax <= q.cp) ? 1.f : 0.f);
        const float g = __fmul_rn(__fmul_rn(gy, sg), inter);     // d out / d lin
        float d;
        if (variant == 3) {
            d = __fmul_rn(__fdiv_rn(g, q.distance), sg);
            s[0] += (double)__fdiv_rn(-g, q.distance);
            s[1] += (double)__fmul_rn(g, __fdiv_rn(-__fsub_rn(ax, q.pp), dsq));
        } else {
            d = __fmul_rn(__fmul_rn(g, q.a), sg);
            s[0] += (double)__fmul_rn(g, ax);   // d/da
            s[1] += (double)g;                  // d/db
        }
        return d;
    }
    __device__ void finalize(const double* s, const float*) {
        if (variant == 3) {
            put_scalar(dp0, __fmul_rn((float)s[0], q.pp), req_p0);
            put_scalar(dp1, __fmul_rn((float)s[1], q.distance), req_p1);
            return;
        }
        const float da = (float)s[0], db = (float)s[1];
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
// Whole-tensor case: 256-bit loads, eight codes packed into one 64-bit store (5 B/element of traffic).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int export_code(float x, float T, const QScale& s, int clip_mode) {
    const float c = quant_code(clip_value(clip_mode, x, T), s);
    return max(-128, min(127, __float2int_rn(c)));
}

__global__ void __launch_bounds__(B2Q_THREADS)
export_int8_flat_kernel(const float* __restrict__ x, int8_t* __restrict__ codes, float* __restrict__ steps, FlatSplit sp,
                        const float* __restrict__ thr, float qlevel, int clip_mode, int fast) {
    b2q_pdl_sync();
    const float T = __ldg(thr);
    const QScale s = make_qscale(T, qlevel, fast != 0);
    const float* xb = x + sp.head;
    int8_t* cb = codes + sp.head;   // head is chosen so that x + head is 32-byte aligned; codes + head must be 8-byte aligned
    const int64_t tile = (int64_t)B2Q_THREADS * 2;
    const int64_t ntiles = (sp.n8 + tile - 1) / tile;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t base = t * tile + threadIdx.x;
        f8 v[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int64_t w = base + (int64_t)k * B2Q_THREADS;
            if (w < sp.n8) v[k] = ld_f8<2>(xb + 8 * w);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int64_t w = base + (int64_t)k * B2Q_THREADS;
            if (w < sp.n8) {
                unsigned long long packed = 0ull;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    packed |= (unsigned long long)(unsigned char)(signed char)export_code(v[k].v[j], T, s, clip_mode) << (8 * j);
                *reinterpret_cast<unsigned long long*>(cb + 8 * w) = packed;
            }
        }
    }
    if (blockIdx.x == 0) {
        const int64_t tid = threadIdx.x;
        int64_t idx = -1;
        if (tid < sp.head) idx = tid;
        else if (tid - sp.head < sp.tail) idx = sp.head + 8 * sp.n8 + (tid - sp.head);
        if (idx >= 0) codes[idx] = (int8_t)export_code(x[idx], T, s, clip_mode);
        if (threadIdx.x == 0 && steps) steps[0] = s.q;
    }
}

__global__ void __launch_bounds__(B2Q_THREADS)
export_int8_kernel(const float* __restrict__ x, int8_t* __restrict__ codes, float* __restrict__ steps, int64_t outer,
                   int64_t groups, int64_t inner, const float* __restrict__ thr, float qlevel, int clip_mode, int fast) {
    const int64_t n = outer * groups * inner;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t g = (groups == 1) ? 0 : (i / inner) % groups;
        const float T = thr[g];
        const QScale s = make_qscale(T, qlevel, fast != 0);
        codes[i] = (int8_t)export_code(x[i], T, s, clip_mode);
        if (i < groups && steps) steps[i] = make_qscale(thr[i], qlevel, false).q;
    }
}

// ------------------------------------------------------------------------------------------------
// Self tests of the two numerical shortcuts (exhaustive over their whole domain; tests/test_gpu_selftest.py)
// ------------------------------------------------------------------------------------------------
__global__ void selftest_div_level_kernel(unsigned long long* bad) {
    // every integer code |c| <= 4 L for L = 2^nbits - 1, nbits = 1..16, and both zeros
    unsigned long long local = 0;
    for (int nbits = 1; nbits <= 16; ++nbits) {
        const float L = (float)((1 << nbits) - 1);
        const LevelDiv d = make_level_div(L);
        const int lim = 4 * ((1 << nbits) - 1);
        for (int c = -lim + (int)(blockIdx.x * blockDim.x + threadIdx.x); c <= lim; c += gridDim.x * blockDim.x) {
            const float cf = (float)c;
            if (__float_as_uint(div_level(cf, d)) != __float_as_uint(__fdiv_rn(cf, L))) ++local;
            if (c == 0 && __float_as_uint(div_level(-0.f, d)) != __float_as_uint(__fdiv_rn(-0.f, L))) ++local;
        }
    }
    if (local) atomicAdd(bad, local);
}

// mode 2: violations of the facts DorefaAbsMax relies on: (a) tanhf(next(x)) >= tanhf(x) for every adjacent pair with
//         both members OUTSIDE the window W, and across W's borders; (b) tanhf(W's lowest float) <= tanhf(x) <=
//         tanhf(first float above W) for x in W; (c) tanhf(-x) == -tanhf(x).  mode 3: largest x (bit pattern) with tanhf(next(x)) < tanhf(x), anywhere;
// mode 4: largest x with tanhf(-x) != -tanhf(x)
__global__ void selftest_tanh_kernel(unsigned long long* bad, int mode) {
    unsigned long long local = 0;
    const unsigned int last = 0x7f7fffffu;
    const float above = tanhf(__uint_as_float(B2Q_TANH_WINDOW_HI + 1u));
    const float bottom = tanhf(__uint_as_float(B2Q_TANH_WINDOW_LO));
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < last;
         b += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((unsigned int)b), xn = __uint_as_float((unsigned int)b + 1u);
        const float t = tanhf(x), tn = tanhf(xn);
        const bool inv = !(tn >= t);
        const bool odd = __float_as_uint(tanhf(-x)) != (__float_as_uint(t) ^ 0x80000000u);
        const bool in_w = b >= B2Q_TANH_WINDOW_LO && b <= B2Q_TANH_WINDOW_HI;
        const bool next_in_w = b + 1 >= B2Q_TANH_WINDOW_LO && b + 1 <= B2Q_TANH_WINDOW_HI;
        if (mode == 2) {
            if (inv && !(in_w && next_in_w)) ++local;          // (a) an inversion may only join two members of W
            if (in_w && !(t <= above && t >= bottom)) ++local; // (b): W's values lie between its bottom and what follows W
            if (odd) ++local;                                  // (c)
        } else if ((mode == 3 && inv) || (mode == 4 && odd)) {
            local = b > local ? b : local;
        }
    }
    if (local) {
        if (mode == 2) atomicAdd(bad, local);
        else atomicMax(bad, local);
    }
}

// mode 5: the integer-pipe float32 -> float64 conversion (f2d_scaled / f2d_abs_scaled, b2q_common.cuh) against the
// conversion instruction over ALL 2^32 bit patterns: equal bits for every finite value (zeros, denormals), `special`
// raised exactly for Inf / NaN; and the scaled value times 2^896 is the value itself.
__global__ void selftest_f2d_kernel(unsigned long long* bad) {
    unsigned long long local = 0;
    for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < (1ull << 32);
         b += (unsigned long long)gridDim.x * blockDim.x) {
        const float x = __uint_as_float((unsigned int)b);
        const bool finite = ((unsigned int)b & 0x7f800000u) != 0x7f800000u;
        bool sp = false, spa = false;
        const double d = f2d_scaled(x, sp) * b2q_two_p896();
        const double da = f2d_abs_scaled(x, spa) * b2q_two_p896();
        const double ref = f2d_cvt(x), refa = f2d_cvt(fabsf(x));
        if (sp == finite || spa == finite) ++local;
        if (finite && (__double_as_longlong(d) != __double_as_longlong(ref) ||
                       __double_as_longlong(da) != __double_as_longlong(refa))) ++local;
        // the scaled image is exact too: scaling the reference down gives the same bits
        bool t = false;
        if (finite && __double_as_longlong(f2d_scaled(x, t)) != __double_as_longlong(ref * b2q_two_m896())) ++local;
    }
    if (local) atomicAdd(bad, local);
}

extern "C" {

int b2q_selftest(b2q_ctx* ctx, int which, int64_t* failures) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(failures != nullptr && which >= 1 && which <= 5,
                "which: 1 level division, 2 tanhf monotonic/odd (count), 3 / 4 where (bit pattern of the largest x), "
                "5 integer-pipe float -> double conversion");
    unsigned long long* d = nullptr;
    B2Q_CHECK_CUDA(cudaMalloc(&d, sizeof(unsigned long long)));
    B2Q_CHECK_CUDA(cudaMemset(d, 0, sizeof(unsigned long long)));
    if (which == 1) selftest_div_level_kernel<<<ctx->num_sms, 256>>>(d);
    else if (which == 5) selftest_f2d_kernel<<<ctx->num_sms * 8, 256>>>(d);
    else selftest_tanh_kernel<<<ctx->num_sms * 8, 256>>>(d, which);
    unsigned long long h = 0;
    cudaError_t e = cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d);
    B2Q_CHECK_CUDA(e);
    *failures = (int64_t)h;
    return 0;
}

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
    b2q_timed_launch tl(ctx, B2Q_KIND_OTHER, 5.0 * (double)n, (cudaStream_t)stream);
    if (groups == 1) {
        FlatSplit sp = b2q_flat_split(x, n);
        if (sp.head <= B2Q_THREADS && (((uintptr_t)(codes + sp.head)) & 7) == 0) {
            const int64_t tile = (int64_t)B2Q_THREADS * 2;
            int64_t grid = (sp.n8 + tile - 1) / tile;
            if (grid < 1) grid = 1;
            b2q_launch(ctx, export_int8_flat_kernel, (unsigned)grid, B2Q_THREADS, (cudaStream_t)stream, x, codes, steps, sp,
                       thr, qlevel, clip_mode, ctx->fast_div);
            B2Q_LAUNCH_CHECK(ctx);
            return 0;
        }
    }
    int64_t grid = (n + B2Q_THREADS - 1) / B2Q_THREADS;
    if (grid > (int64_t)ctx->num_sms * 32) grid = (int64_t)ctx->num_sms * 32;
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
    b2q_slot* slot = b2q_take_slot(ctx, st);
    UpdateArgs u;
    memset(&u, 0, sizeof(u));
    u.stat_out = slot->scale;
    const int64_t groups = per_channel ? rows : 1, inner = per_channel ? cols : rows * cols;
    int rc = launch_reduce<true>(ctx, slot, x, 1, groups, inner, kNoPrescale, u, st);
    if (rc) return rc;
    const int add = req == B2Q_REQ_ADD;
    if (per_channel) {
        WnqFwd<true> op = {x, y, slot->scale, cols, qlevel, add, 0.f, {}};
        return launch_ew(ctx, op, rows * cols, 8.0, st);
    }
    WnqFwd<false> op = {x, y, slot->scale, cols, qlevel, add, 0.f, {}};
    return launch_ew(ctx, op, rows * cols, 8.0, st);
}

int b2q_wnq_bwd_f32(b2q_ctx* ctx, const float* x, const float* dy, float* dx, int64_t rows, int64_t cols,
                    int per_channel, int req, void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(x && dy && dx && rows >= 1 && cols >= 1, "bad argument");
    if (req == B2Q_REQ_NULL) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    b2q_slot* slot = b2q_take_slot(ctx, st);
    UpdateArgs u;
    memset(&u, 0, sizeof(u));
    u.stat_out = slot->scale;
    const int64_t groups = per_channel ? rows : 1, inner = per_channel ? cols : rows * cols;
    int rc = launch_reduce<true>(ctx, slot, x, 1, groups, inner, kNoPrescale, u, st);
    if (rc) return rc;
    SegPlan pl = b2q_seg_plan(x, dy, 1, groups, inner, ctx->num_sms * 8);
    {
        b2q_timed_launch tl(ctx, B2Q_KIND_OTHER, 8.0 * (double)(rows * cols), st);
        wnq_bwd_sum_kernel<<<(unsigned)(groups * pl.S * pl.P), 128, 0, st>>>(x, dy, pl, slot->scale, slot->clip, slot);
        B2Q_LAUNCH_CHECK(ctx);
    }
    const int add = req == B2Q_REQ_ADD;
    if (per_channel) {
        WnqBwdApply<true> op = {x, dy, dx, slot->scale, slot->clip, cols, add, 0.f, 0.f};
        return launch_ew(ctx, op, rows * cols, 12.0, st);
    }
    WnqBwdApply<false> op = {x, dy, dx, slot->scale, slot->clip, cols, add, 0.f, 0.f};
    return launch_ew(ctx, op, rows * cols, 12.0, st);
}

int b2q_pact_bwd_f32(b2q_ctx* ctx, const float* x, const float* dy, float* dx, float* dgamma, const float* gamma,
                     int64_t n, int two_sided, int req, int req_gamma, void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(x && dy && gamma && n >= 1, "bad argument");
    B2Q_REQUIRE(req == B2Q_REQ_NULL || dx, "null dx");
    PactBwd op = {x, dy, req == B2Q_REQ_NULL ? nullptr : dx, dgamma, gamma, two_sided, req == B2Q_REQ_ADD, req_gamma, 0.f};
    return launch_ew(ctx, op, n, 12.0, (cudaStream_t)stream);
}

int b2q_dorefa_fwd_f32(b2q_ctx* ctx, const float* x, float* y, float* vmax_out, int64_t n, float qlevel, int req,
                       void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(x && y && vmax_out && n >= 1, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (ctx->dorefa_tanh_max) {   // element-wise max of |tanh(w)| (PACT.py:48 as written)
        DorefaMax mop = {x, vmax_out};
        rc = launch_ew(ctx, mop, n, 4.0, st);
    } else {                      // max|w| + the window around tanhf's non-monotonic step: the same float, no tanh pass
        DorefaAbsMax mop = {x, vmax_out};
        rc = launch_ew(ctx, mop, n, 4.0, st);
    }
    if (rc || req == B2Q_REQ_NULL) return rc;
    DorefaFwd op = {x, y, vmax_out, qlevel, req == B2Q_REQ_ADD, 0.f, {}};
    return launch_ew(ctx, op, n, 8.0, st);
}

int b2q_dorefa_bwd_f32(b2q_ctx* ctx, const float* x, const float* dy, float* dx, const float* vmax, int64_t n, int req,
                       void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(x && dy && dx && vmax && n >= 1, "bad argument");
    if (req == B2Q_REQ_NULL) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    b2q_slot* slot = b2q_take_slot(ctx, st);
    DorefaBwdSum sop = {x, dy, vmax, slot->scale, 0.f, 0.f};
    int rc = launch_ew(ctx, sop, n, 8.0, st);
    if (rc) return rc;
    DorefaBwdApply op = {x, dy, dx, vmax, slot->scale, req == B2Q_REQ_ADD, 0.f, 0.f, 0.f};
    return launch_ew(ctx, op, n, 12.0, st);
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
    if (variant == 3) {
        QilFwd<1> op = {variant, x, y, p0, p1, qlevel, req == B2Q_REQ_ADD, {}, {}};
        return launch_ew(ctx, op, n, 8.0, st);
    }
    QilFwd<0> op = {variant, x, y, p0, p1, qlevel, req == B2Q_REQ_ADD, {}, {}};
    return launch_ew(ctx, op, n, 8.0, st);
}

int b2q_qil_bwd_f32(b2q_ctx* ctx, int variant, const float* x, const float* dy, float* dx, const float* p0,
                    const float* p1, float* dp0, float* dp1, int64_t n, int req, int req_p0, int req_p1, void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(variant >= 1 && variant <= 3, "variant must be 1, 2 or 3");
    B2Q_REQUIRE(x && dy && p0 && p1 && n >= 1, "bad argument");
    B2Q_REQUIRE(req == B2Q_REQ_NULL || dx, "null dx");
    QilBwd op = {variant, x, dy, req == B2Q_REQ_NULL ? nullptr : dx, p0, p1, dp0, dp1, req == B2Q_REQ_ADD, req_p0, req_p1,
                 {}, 0.f};
    return launch_ew(ctx, op, n, 12.0, (cudaStream_t)stream);
}

}  // extern "C"
