// K4/K8: the single-sweep clip + quantize-dequantize kernels, and K5/K6 backward passes.
#pragma once
#include "b2q_common.cuh"
#include "b2q_reduce.cuh"

// Compile-time tuning of the flat sweeps (tools/sweep.cu on B200 picks these; see profiles/).
#ifndef B2Q_QDQ_UNROLL
#define B2Q_QDQ_UNROLL 2
#endif
#ifndef B2Q_QDQ_LDPOL
#define B2Q_QDQ_LDPOL 2   // x: read-only path, L2 evict_first (last forward use of the line)
#endif
#ifndef B2Q_QDQ_STPOL
#define B2Q_QDQ_STPOL 0   // y: plain store -- the convolution reads it next, let it live in L2
#endif
#ifndef B2Q_STREAM_BYTES
#define B2Q_STREAM_BYTES (96ll << 20)   // outputs larger than this cannot stay in L2 anyway: store them evict_first
#endif
#ifndef B2Q_BWD_UNROLL
#define B2Q_BWD_UNROLL 2
#endif
#ifndef B2Q_BWD_LDPOL
#define B2Q_BWD_LDPOL 2
#endif
#ifndef B2Q_BWD_STPOL
#define B2Q_BWD_STPOL 0
#endif

// ------------------------------------------------------------------------------------------------
// Exact emulation of the reference's three kernels   t = x / q ; c = round(t) ; y = c * q
// (symbol/quant_ops.py:28,40) = fl(roundf(fl(x/q)) * q).
//
// IEEE division (~9 issue slots) plus roundf (~4) would make the sweep issue-bound at HBM speed, so the
// common case uses the reciprocal: t' = fl(x * r), r = fl(1/q).  Both roundings are within 2^-24 relative, and
// so is fl(x/q), hence |t' - fl(x/q)| <= 0.75 * 2^-22 * |t'| and rint(t') == roundf(fl(x/q)) whenever t' is
// farther than |t'| * 2^-21 from the nearest half-integer (margin 2.6x).  Everything else -- exact ties,
// near-ties, codes >= 2^20, NaN/Inf, q outside [2^-100, 2^100] -- takes the reference arithmetic verbatim.
// The result is bit-identical to the slow path for every input; only how often the slow path runs
// (~1e-4 of elements for |code| ~ 100) depends on the data.
// ------------------------------------------------------------------------------------------------
struct QScale {
    float q;   // quantisation step fl(T / L)
    float r;   // fl(1/q), or NaN when the fast path must not be used
};

__device__ __forceinline__ QScale make_qscale(float T, float qlevel, bool fast) {
    QScale s;
    s.q = qlevel > 0.f ? __fdiv_rn(T, qlevel) : T;
    const float aq = fabsf(s.q);
    s.r = (fast && aq > 0x1p-100f && aq < 0x1p100f) ? __frcp_rn(s.q) : __int_as_float(0x7fc00000);
    return s;
}

// reference arithmetic, verbatim
__device__ __forceinline__ float quant_code_exact(float x, float q) { return roundf(__fdiv_rn(x, q)); }

// fast candidate + safety test.  safe <=> |t - rint(t)| + |t| * 2^-21 < 0.5  (false for NaN and |t| >= 2^20)
__device__ __forceinline__ float quant_code_try(float x, float r, bool& safe) {
    const float t = __fmul_rn(x, r);
    const float c = rintf(t);
    const float d = __fsub_rn(t, c);
    safe = safe && (__fmaf_rn(fabsf(t), 0x1p-21f, fabsf(d)) < 0.5f);
    return c;
}

__device__ __forceinline__ float quant_code(float x, const QScale& s) {
    bool safe = true;
    float c = quant_code_try(x, s.r, safe);
    if (!safe) c = quant_code_exact(x, s.q);
    return c;
}

__device__ __forceinline__ float clip_value(int clip, float x, float T) {
    switch (clip) {
        case B2Q_CLIP_SYM: return mx_clip(x, -T, T);
        case B2Q_CLIP_WHERE_LE: return (fabsf(x) <= T) ? x : __fmul_rn(T, mx_sign(x));
        case B2Q_CLIP_ZERO_T: return mx_clip(x, 0.f, T);
        case B2Q_CLIP_PACT: return (x < T) ? x : T;
        case B2Q_CLIP_WHERE_LT: return (fabsf(x) < T) ? x : __fmul_rn(T, mx_sign(x));
        default: return x;
    }
}

// Clip on the fast path: min/max forms that equal the reference's comparison-based clips for non-NaN inputs and a
// non-negative threshold.  The NaN-PROPAGATING min/max are used (fminf/fmaxf would turn a NaN input into -T or 0), so a
// NaN input reaches the safety test, fails it, and the word is redone with clip_value + the reference arithmetic
// (mx.nd.clip passes NaN through; where(|x|<=a, x, a*sign(x)) gives a*0).  The caller disables the fast path when the
// threshold is negative.
template <int CLIP>
__device__ __forceinline__ float clip_fast(float x, float T) {
    if (CLIP == B2Q_CLIP_SYM || CLIP == B2Q_CLIP_WHERE_LE || CLIP == B2Q_CLIP_WHERE_LT) return fmin_nan(fmax_nan(x, -T), T);
    if (CLIP == B2Q_CLIP_ZERO_T) return fmin_nan(fmax_nan(x, 0.f), T);
    if (CLIP == B2Q_CLIP_PACT) return fmin_nan(x, T);
    return x;
}

// Eight elements at once: one divergence point per 256-bit word instead of one per element.
template <int CLIP>
__device__ __forceinline__ void qdq8(const f8& in, f8& out, float Tc, const QScale& s) {
    bool safe = true;
#pragma unroll
    for (int j = 0; j < 8; ++j) out.v[j] = __fmul_rn(quant_code_try(clip_fast<CLIP>(in.v[j], Tc), s.r, safe), s.q);
    if (!safe) {
#pragma unroll
        for (int j = 0; j < 8; ++j) out.v[j] = __fmul_rn(quant_code_exact(clip_value(CLIP, in.v[j], Tc), s.q), s.q);
    }
}

__device__ __forceinline__ int32_t code_to_i32(float c) {
    return __float2int_rn(c);  // saturating, NaN -> 0
}

struct QdqArgs {
    const float* thr;       // [groups] scale threshold (device) or null -> thr_imm
    const float* clip_thr;  // [groups] clip threshold (device) or null -> same as thr / clip_imm
    float thr_imm, clip_imm;
    float qlevel;
    int fast;
    int32_t* codes;
    int clip_mode, do_round, req;   // used by the generic kernels only
};

__device__ __forceinline__ float qdq_generic(const QdqArgs& a, float x, float Tc, const QScale& s, float& code) {
    const float c = clip_value(a.clip_mode, x, Tc);
    if (!a.do_round) { code = 0.f; return c; }
    code = quant_code(c, s);
    return __fmul_rn(code, s.q);
}

// ------------------------------------------------------------------------------------------------
// HOT kernel: whole tensor, clip none / symmetric, round, req=write, no codes.
// Tiles are visited in DESCENDING address order: the reduction that precedes it walked ascending, so the
// tail of x is still in the 126 MB L2.  x is loaded evict-first (last use); y is stored plainly because the
// convolution reads it next.
// ------------------------------------------------------------------------------------------------
template <int CLIP, int UNROLL, int LDPOL, int STPOL, bool DEFERRED>
__global__ void __launch_bounds__(B2Q_THREADS)
qdq_flat_hot_kernel(const float* __restrict__ x, float* __restrict__ y, FlatSplit sp, QdqArgs a, int reverse,
                    DeferredUpdate d, int clip_with_fresh) {
    const float* xb = x + sp.head;
    float* yb = y + sp.head;
    const int64_t tile = (int64_t)B2Q_THREADS * UNROLL;
    const int64_t ntiles = (sp.n8 + tile - 1) / tile;
    f8 v[UNROLL];
    if (DEFERRED) {
        // The fused forward launches this kernel right behind its reduction, which only READS x and has itself waited
        // for whatever produced x before letting us start: the first tile can be requested before the dependency wait,
        // so its HBM latency overlaps the reduction's last wave.  (Only x: the statistic and y come after the wait.)
        const int64_t t0 = reverse ? (ntiles - 1 - (int64_t)blockIdx.x) : (int64_t)blockIdx.x;
        const int64_t base0 = t0 * tile + threadIdx.x;
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) {
            const int64_t i = base0 + (int64_t)k * B2Q_THREADS;
            if ((int64_t)blockIdx.x < ntiles && i < sp.n8) v[k] = ld_f8<LDPOL>(xb + 8 * i);
        }
    }
    b2q_pdl_sync();
    float T, Tc;
    if (DEFERRED) {
        float stat;
        if (d.is_max) {
            // one epoch-tagged word, read by every thread (broadcast L2 hit): no shared memory, no barrier
            const unsigned long long word = *d.max64;
            stat = __uint_as_float((unsigned int)(word & 0xffffffffull));
            if (blockIdx.x == 0 && threadIdx.x == 0) *d.epoch = (unsigned int)(word >> 32);   // consume the tag
        } else {
            // sums: every block combines the double partials in the same fixed order
            __shared__ double smem[32];
            __shared__ float s_stat;
            double acc = 0.0;
            for (int i = threadIdx.x; i < d.n_partials; i += blockDim.x) acc += d.partial[i];
            const double tot = block_reduce<false>(acc, smem);
            if (threadIdx.x == 0) s_stat = __fdiv_rn((float)tot, d.count);
            __syncthreads();
            stat = s_stat;
        }
        const float a_old = d.u.aux ? d.aux_old[0] : 0.f;
        float fresh, next;
        compute_update(d.u.mode, d.u.p0, d.u.p1, a_old, stat, fresh, next);
        const float after = d.u.write_aux ? next : a_old;
        T = d.u.use_aux_as_scale ? after : fresh;
        Tc = clip_with_fresh ? fresh : T;
        if (blockIdx.x == 0 && threadIdx.x == 0 && d.u.write_aux && d.u.aux) d.u.aux[0] = next;
    } else {
        T = a.thr ? __ldg(a.thr) : a.thr_imm;
        Tc = a.clip_thr ? __ldg(a.clip_thr) : (a.thr ? T : a.clip_imm);
    }
    const bool needs_pos = (CLIP != B2Q_CLIP_NONE && CLIP != B2Q_CLIP_PACT);
    const QScale s = make_qscale(T, a.qlevel, a.fast != 0 && !(needs_pos && !(Tc >= 0.f)));
    for (int64_t tt = blockIdx.x; tt < ntiles; tt += gridDim.x) {
        const int64_t t = reverse ? (ntiles - 1 - tt) : tt;
        const int64_t base = t * tile + threadIdx.x;
        if (!DEFERRED || tt != (int64_t)blockIdx.x) {
#pragma unroll
            for (int k = 0; k < UNROLL; ++k) {
                const int64_t i = base + (int64_t)k * B2Q_THREADS;
                if (i < sp.n8) v[k] = ld_f8<LDPOL>(xb + 8 * i);
            }
        }
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) {
            const int64_t i = base + (int64_t)k * B2Q_THREADS;
            if (i < sp.n8) {
                f8 o;
                qdq8<CLIP>(v[k], o, Tc, s);
                st_f8<STPOL>(yb + 8 * i, o);
            }
        }
    }
    if (blockIdx.x == 0) {  // unaligned head / tail scalars
        const int64_t tid = threadIdx.x;
        int64_t idx = -1;
        if (tid < sp.head) idx = tid;
        else if (tid - sp.head < sp.tail) idx = sp.head + 8 * sp.n8 + (tid - sp.head);
        if (idx >= 0) {
            y[idx] = __fmul_rn(quant_code_exact(clip_value(CLIP, x[idx], Tc), s.q), s.q);
        }
    }
}

// Generic whole-tensor sweep: any clip mode, optional rounding, req add, optional codes (runtime flags).
template <int UNROLL>
__global__ void __launch_bounds__(B2Q_THREADS)
qdq_flat_generic_kernel(const float* __restrict__ x, float* __restrict__ y, FlatSplit sp, QdqArgs a) {
    b2q_pdl_sync();
    const float T = a.thr ? __ldg(a.thr) : a.thr_imm;
    const float Tc = a.clip_thr ? __ldg(a.clip_thr) : (a.thr ? T : a.clip_imm);
    const QScale s = make_qscale(T, a.qlevel, a.fast != 0);
    const float* xb = x + sp.head;
    float* yb = y + sp.head;
    const bool add = (a.req == B2Q_REQ_ADD);
    const int64_t tile = (int64_t)B2Q_THREADS * UNROLL;
    const int64_t ntiles = (sp.n8 + tile - 1) / tile;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t base = t * tile + threadIdx.x;
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) {
            const int64_t i = base + (int64_t)k * B2Q_THREADS;
            if (i < sp.n8) {
                const f8 v = ld_f8<1>(xb + 8 * i);
                f8 o;
                int cc[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float code;
                    o.v[j] = qdq_generic(a, v.v[j], Tc, s, code);
                    cc[j] = code_to_i32(code);
                }
                if (add) {
                    const f8 old = ld_f8<0>(yb + 8 * i);
#pragma unroll
                    for (int j = 0; j < 8; ++j) o.v[j] = __fadd_rn(old.v[j], o.v[j]);
                }
                st_f8<0>(yb + 8 * i, o);
                if (a.codes) st_i8(a.codes + sp.head + 8 * i, cc);
            }
        }
    }
    if (blockIdx.x == 0) {
        const int64_t tid = threadIdx.x;
        int64_t idx = -1;
        if (tid < sp.head) idx = tid;
        else if (tid - sp.head < sp.tail) idx = sp.head + 8 * sp.n8 + (tid - sp.head);
        if (idx >= 0) {
            float code;
            float o = qdq_generic(a, x[idx], Tc, s, code);
            if (add) o = __fadd_rn(y[idx], o);
            y[idx] = o;
            if (a.codes) a.codes[idx] = code_to_i32(code);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Segmented sweep over (outer, groups, inner) with one threshold per group and optional fold-BN prescale.
// Also writes the folded bias when `bias` is given (fold_bn_v1_gdrq.py:113).
// ------------------------------------------------------------------------------------------------
struct FoldBias {
    float* bias;         // [rows] or null
    const float* beta;
    const float* mean;
};

// Grouped ACTIVATIONS (GDRQ group_size > 0 on NCHW, tensors of hundreds of MB): every piece is a contiguous range with
// one threshold, so the block runs the hot sweep's inner loop (256-bit accesses, 8-wide fast path) on its range.
template <int CLIP>
__global__ void __launch_bounds__(B2Q_THREADS)
qdq_seg_hot_kernel(const float* __restrict__ x, float* __restrict__ y, SegPlan pl, QdqArgs a) {
    b2q_pdl_sync();
    const SegPiece pc = seg_piece(pl);
    const float T = __ldg(a.thr + pc.g);
    const float Tc = a.clip_thr ? __ldg(a.clip_thr + pc.g) : T;
    const bool needs_pos = (CLIP != B2Q_CLIP_NONE && CLIP != B2Q_CLIP_PACT);
    const QScale s = make_qscale(T, a.qlevel, a.fast != 0 && !(needs_pos && !(Tc >= 0.f)));
    // the piece's (row, word) space flattened into one index, two 256-bit words in flight per thread
    const unsigned wpr = (unsigned)((pc.i1 - pc.i0) >> 3);
    const unsigned total = (unsigned)(pc.o1 - pc.o0) * wpr;
    for (unsigned w0 = threadIdx.x; w0 < total; w0 += 2 * blockDim.x) {
        const unsigned w1 = w0 + blockDim.x;
        const unsigned o0 = w0 / wpr, i0 = w0 - o0 * wpr;
        const int64_t off0 = ((pc.o0 + o0) * pl.groups + pc.g) * pl.inner + pc.i0 + 8 * (int64_t)i0;
        int64_t off1 = 0;
        f8 v0, v1;
        v0 = ld_f8<B2Q_QDQ_LDPOL>(x + off0);
        if (w1 < total) {
            const unsigned o1 = w1 / wpr, i1 = w1 - o1 * wpr;
            off1 = ((pc.o0 + o1) * pl.groups + pc.g) * pl.inner + pc.i0 + 8 * (int64_t)i1;
            v1 = ld_f8<B2Q_QDQ_LDPOL>(x + off1);
        }
        f8 r;
        qdq8<CLIP>(v0, r, Tc, s);
        st_f8<1>(y + off0, r);
        if (w1 < total) {
            qdq8<CLIP>(v1, r, Tc, s);
            st_f8<1>(y + off1, r);
        }
    }
}

template <int VEC>
__global__ void __launch_bounds__(128)
qdq_seg_kernel(const float* __restrict__ x, float* __restrict__ y, SegPlan pl, Prescale ps, FoldBias fb, QdqArgs a) {
    b2q_pdl_sync();
    const SegPiece pc = seg_piece(pl);
    const float T = a.thr ? __ldg(a.thr + pc.g) : a.thr_imm;
    const float Tc = a.clip_thr ? __ldg(a.clip_thr + pc.g) : (a.thr ? T : a.clip_imm);
    const QScale qs = make_qscale(T, a.qlevel, a.fast != 0);
    const bool add = (a.req == B2Q_REQ_ADD);
    for (int64_t o = pc.o0; o < pc.o1; ++o) {
        const int64_t row = o * pl.groups + pc.g;
        const float* xb = x + row * pl.inner;
        float* yb = y + row * pl.inner;
        int32_t* cb = a.codes ? a.codes + row * pl.inner : nullptr;
        const float f = ps.gamma ? prescale_factor(ps, row) : 1.f;
        if (fb.bias && pc.p == 0 && threadIdx.x == 0) {
            // bias = beta - mean*gamma/sqrt(var+eps), each step rounded (fold_bn_v1_gdrq.py:113)
            const float den = __fsqrt_rn(__fadd_rn(ps.var[row], ps.eps));
            fb.bias[row] = __fsub_rn(fb.beta[row], __fdiv_rn(__fmul_rn(fb.mean[row], ps.gamma[row]), den));
        }
        if (VEC == 4) {
            const float4* x4 = reinterpret_cast<const float4*>(xb);
            float4* y4 = reinterpret_cast<float4*>(yb);
            const int64_t end = pc.i1 >> 2;
            for (int64_t i0 = (pc.i0 >> 2) + threadIdx.x; i0 < end; i0 += 4 * (int64_t)blockDim.x) {
                float4 vv[4];   // four independent 128-bit loads in flight per thread
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int64_t j = i0 + (int64_t)k * blockDim.x;
                    if (j < end) vv[k] = x4[j];
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int64_t i = i0 + (int64_t)k * blockDim.x;
                    if (i >= end) break;
                    float4 v = vv[k];
                    if (ps.gamma) { v.x = __fmul_rn(v.x, f); v.y = __fmul_rn(v.y, f); v.z = __fmul_rn(v.z, f); v.w = __fmul_rn(v.w, f); }
                    float4 r;
                    float cx, cy, cz, cw;
                    r.x = qdq_generic(a, v.x, Tc, qs, cx);
                    r.y = qdq_generic(a, v.y, Tc, qs, cy);
                    r.z = qdq_generic(a, v.z, Tc, qs, cz);
                    r.w = qdq_generic(a, v.w, Tc, qs, cw);
                    if (add) {
                        const float4 old = y4[i];
                        r.x = __fadd_rn(old.x, r.x); r.y = __fadd_rn(old.y, r.y);
                        r.z = __fadd_rn(old.z, r.z); r.w = __fadd_rn(old.w, r.w);
                    }
                    y4[i] = r;
                    if (cb) {
                        cb[4 * i + 0] = code_to_i32(cx); cb[4 * i + 1] = code_to_i32(cy);
                        cb[4 * i + 2] = code_to_i32(cz); cb[4 * i + 3] = code_to_i32(cw);
                    }
                }
            }
        } else {
            for (int64_t i = pc.i0 + threadIdx.x; i < pc.i1; i += blockDim.x) {
                float v = xb[i];
                if (ps.gamma) v = __fmul_rn(v, f);
                float c;
                float r = qdq_generic(a, v, Tc, qs, c);
                if (add) r = __fadd_rn(yb[i], r);
                yb[i] = r;
                if (cb) cb[i] = code_to_i32(c);
            }
        }
    }
}

// Grouped ACTIVATIONS with SHORT rows that are not a multiple of eight floats (GDRQ group_size on 14x14 / 7x7 maps: rows
// of 196 / 49 floats): the row-at-a-time loops above keep 49 of 128 threads busy with one access each.  Here the piece's
// (row, element) space is flattened into one index walked incrementally (no division per element), U accesses of V
// floats in flight per thread.  V = 4: rows a multiple of four floats and 16-byte aligned; V = 1: anything.
struct FlatWalk {
    unsigned len, qT, rT, o, i;
    unsigned long long total;
    int64_t ostride, base;      // element offset of the piece's first row / between rows
};

__device__ __forceinline__ FlatWalk flat_walk(const SegPlan& pl, const SegPiece& pc, int V) {
    FlatWalk w;
    w.len = (unsigned)((pc.i1 - pc.i0) / V);
    w.total = (unsigned long long)(pc.o1 - pc.o0) * w.len;
    w.base = (pc.o0 * pl.groups + pc.g) * pl.inner + pc.i0;
    w.ostride = pl.groups * pl.inner;
    w.qT = blockDim.x / w.len;
    w.rT = blockDim.x % w.len;
    w.o = threadIdx.x / w.len;
    w.i = threadIdx.x % w.len;
    return w;
}

__device__ __forceinline__ void flat_walk_next(FlatWalk& w) {
    w.i += w.rT; w.o += w.qT;
    if (w.i >= w.len) { w.i -= w.len; ++w.o; }
}

template <int V>
__global__ void __launch_bounds__(B2Q_THREADS)
qdq_seg_flatidx_kernel(const float* __restrict__ x, float* __restrict__ y, SegPlan pl, QdqArgs a) {
    b2q_pdl_sync();
    constexpr int U = (V == 4) ? 4 : 8;
    const SegPiece pc = seg_piece(pl);
    const float T = a.thr ? __ldg(a.thr + pc.g) : a.thr_imm;
    const float Tc = a.clip_thr ? __ldg(a.clip_thr + pc.g) : (a.thr ? T : a.clip_imm);
    const QScale qs = make_qscale(T, a.qlevel, a.fast != 0);
    const bool add = (a.req == B2Q_REQ_ADD);
    FlatWalk w = flat_walk(pl, pc, V);
    for (unsigned long long t = threadIdx.x; t < w.total; t += (unsigned long long)U * blockDim.x) {
        float v[U][V];
        int64_t off[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            off[k] = w.base + (int64_t)w.o * w.ostride + (int64_t)w.i * V;
            if (t + (unsigned long long)k * blockDim.x < w.total) {
                if (V == 4) {
                    const float4 t4 = *reinterpret_cast<const float4*>(x + off[k]);
                    v[k][0] = t4.x; v[k][V > 1 ? 1 : 0] = t4.y; v[k][V > 2 ? 2 : 0] = t4.z; v[k][V > 3 ? 3 : 0] = t4.w;
                } else {
                    v[k][0] = x[off[k]];
                }
            } else {
                off[k] = -1;
            }
            flat_walk_next(w);
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            if (off[k] < 0) continue;
            float r[V];
#pragma unroll
            for (int e = 0; e < V; ++e) {
                float code;
                r[e] = qdq_generic(a, v[k][e], Tc, qs, code);
                if (add) r[e] = __fadd_rn(y[off[k] + e], r[e]);
            }
            if (V == 4) *reinterpret_cast<float4*>(y + off[k]) = make_float4(r[0], r[V > 1 ? 1 : 0], r[V > 2 ? 2 : 0], r[V > 3 ? 3 : 0]);
            else y[off[k]] = r[0];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Row-fused weight path (K2 / K9): one WARP per row (out-channel or GDRQ group), ONE kernel.
// pass 1: statistic of the row (max|w| or sum|w|, optional fold-BN prescale) -> warp shuffle tree (fixed order);
// threshold update in registers (lane 0 writes aux); pass 2: clip + QDQ of the same row, which is still in L1.
// Replaces reduce + (ticket / last block) + sweep for per-channel weights: depthwise 3x3 rows have 9 elements.
// ------------------------------------------------------------------------------------------------
#define B2Q_ROWS_FUSED_MAX_INNER 16384

template <bool IS_MAX>
__global__ void __launch_bounds__(B2Q_THREADS)
rows_fused_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t rows, int64_t inner, Prescale ps,
                  FoldBias fb, UpdateArgs u, QdqArgs a, int clip_with_fresh) {
    b2q_pdl_sync();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * nw + wid;
    if (row >= rows) return;
    const float* xb = x + row * inner;
    float* yb = y + row * inner;
    const float f = ps.gamma ? prescale_factor(ps, row) : 1.f;
    const bool vec = ((((uintptr_t)xb) & 15) == 0) && ((((uintptr_t)yb) & 15) == 0) && ((inner & 3) == 0);
    double acc = 0.0;
    float m = 0.f;
    if (vec) {
        const float4* x4 = reinterpret_cast<const float4*>(xb);
        for (int64_t i = lane; i < (inner >> 2); i += 32) {
            float4 v = x4[i];
            if (ps.gamma) { v.x = __fmul_rn(v.x, f); v.y = __fmul_rn(v.y, f); v.z = __fmul_rn(v.z, f); v.w = __fmul_rn(v.w, f); }
            acc1<IS_MAX>(acc, m, v.x); acc1<IS_MAX>(acc, m, v.y); acc1<IS_MAX>(acc, m, v.z); acc1<IS_MAX>(acc, m, v.w);
        }
    } else {
        for (int64_t i = lane; i < inner; i += 32) {
            float v = xb[i];
            if (ps.gamma) v = __fmul_rn(v, f);
            acc1<IS_MAX>(acc, m, v);
        }
    }
    float stat;
    if (IS_MAX) stat = warp_max(m);
    else stat = __fdiv_rn((float)warp_sum(acc), (float)inner);
    const float a_old = u.aux ? u.aux[row] : 0.f;
    float fresh, next;
    compute_update(u.mode, u.p0, u.p1, a_old, stat, fresh, next);
    __syncwarp();
    if (lane == 0 && u.write_aux && u.aux) u.aux[row] = next;
    const float after = u.write_aux ? next : a_old;
    const float T = u.use_aux_as_scale ? after : fresh;
    const float Tc = clip_with_fresh ? fresh : T;
    if (fb.bias && lane == 0) {
        const float den = __fsqrt_rn(__fadd_rn(ps.var[row], ps.eps));
        fb.bias[row] = __fsub_rn(fb.beta[row], __fdiv_rn(__fmul_rn(fb.mean[row], ps.gamma[row]), den));
    }
    if (a.req == B2Q_REQ_NULL) return;
    const QScale qs = make_qscale(T, a.qlevel, a.fast != 0);
    const bool add = (a.req == B2Q_REQ_ADD);
    if (vec) {
        const float4* x4 = reinterpret_cast<const float4*>(xb);
        float4* y4 = reinterpret_cast<float4*>(yb);
        for (int64_t i = lane; i < (inner >> 2); i += 32) {
            float4 v = x4[i];
            if (ps.gamma) { v.x = __fmul_rn(v.x, f); v.y = __fmul_rn(v.y, f); v.z = __fmul_rn(v.z, f); v.w = __fmul_rn(v.w, f); }
            float4 r;
            float c;
            r.x = qdq_generic(a, v.x, Tc, qs, c); r.y = qdq_generic(a, v.y, Tc, qs, c);
            r.z = qdq_generic(a, v.z, Tc, qs, c); r.w = qdq_generic(a, v.w, Tc, qs, c);
            if (add) {
                const float4 old = y4[i];
                r.x = __fadd_rn(old.x, r.x); r.y = __fadd_rn(old.y, r.y); r.z = __fadd_rn(old.z, r.z); r.w = __fadd_rn(old.w, r.w);
            }
            y4[i] = r;
        }
    } else {
        for (int64_t i = lane; i < inner; i += 32) {
            float v = xb[i];
            if (ps.gamma) v = __fmul_rn(v, f);
            float c;
            float r = qdq_generic(a, v, Tc, qs, c);
            if (add) r = __fadd_rn(yb[i], r);
            yb[i] = r;
        }
    }
}

// Long rows (1024 <= inner <= B2Q_ROWS_CTA_MAX_INNER, e.g. 512x512x3x3: 4608 per out-channel): one CTA per row.  The
// row is staged ONCE in shared memory -- by a single bulk asynchronous copy (TMA engine, mbarrier completion) when it is
// 16-byte aligned, by cooperative loads otherwise -- and both passes read it from there, so every weight byte crosses
// L2 once and a 512-row tensor spreads over all SMs instead of 64 eight-warp blocks streaming 18 KB per warp twice.
#define B2Q_ROWS_CTA_MIN_INNER 1024
#define B2Q_ROWS_CTA_MAX_INNER 49152   // 192 KB of dynamic shared memory (opt-in above 48 KB)

template <bool IS_MAX>
__global__ void __launch_bounds__(B2Q_THREADS)
rows_cta_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t rows, int64_t inner, Prescale ps,
                FoldBias fb, UpdateArgs u, QdqArgs a, int clip_with_fresh) {
    extern __shared__ __align__(128) float s_row[];
    __shared__ double s_red[32];
    __shared__ float s_stat;
    __shared__ __align__(8) unsigned long long s_bar;
    b2q_pdl_sync();
    const int64_t row = blockIdx.x;
    const float* xb = x + row * inner;
    float* yb = y + row * inner;
    const int n = (int)inner;
    const bool bulk = ((((uintptr_t)xb) & 15) == 0) && ((n & 3) == 0);
    if (bulk) {
        if (threadIdx.x == 0) {
            mbar_init(&s_bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            mbar_expect_tx(&s_bar, (unsigned)n * 4u);
            bulk_g2s(s_row, xb, (unsigned)n * 4u, &s_bar);
        }
        mbar_wait(&s_bar, 0);
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) s_row[i] = xb[i];
        __syncthreads();
    }
    const float f = ps.gamma ? prescale_factor(ps, row) : 1.f;
    double acc = 0.0;
    float m = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {   // consecutive threads, consecutive words: conflict free
        float v = s_row[i];
        if (ps.gamma) {
            v = __fmul_rn(v, f);
            s_row[i] = v;                                  // the second pass reads the folded weight
        }
        acc1<IS_MAX>(acc, m, v);
    }
    const double tot = block_reduce<IS_MAX>(IS_MAX ? (double)m : acc, s_red);
    if (threadIdx.x == 0) s_stat = IS_MAX ? (float)tot : __fdiv_rn((float)tot, (float)inner);
    __syncthreads();
    const float stat = s_stat;
    const float a_old = u.aux ? u.aux[row] : 0.f;
    float fresh, next;
    compute_update(u.mode, u.p0, u.p1, a_old, stat, fresh, next);
    __syncthreads();   // every thread has read aux[row] before thread 0 overwrites it
    if (threadIdx.x == 0) {
        if (u.write_aux && u.aux) u.aux[row] = next;
        if (fb.bias) {
            const float den = __fsqrt_rn(__fadd_rn(ps.var[row], ps.eps));
            fb.bias[row] = __fsub_rn(fb.beta[row], __fdiv_rn(__fmul_rn(fb.mean[row], ps.gamma[row]), den));
        }
    }
    if (a.req == B2Q_REQ_NULL) return;
    const float after = u.write_aux ? next : a_old;
    const float T = u.use_aux_as_scale ? after : fresh;
    const float Tc = clip_with_fresh ? fresh : T;
    const QScale qs = make_qscale(T, a.qlevel, a.fast != 0);
    const bool add = (a.req == B2Q_REQ_ADD);
    const bool vec_out = ((((uintptr_t)yb) & 15) == 0) && ((n & 3) == 0);
    if (vec_out) {
        const float4* s4 = reinterpret_cast<const float4*>(s_row);
        float4* y4 = reinterpret_cast<float4*>(yb);
        for (int i = threadIdx.x; i < (n >> 2); i += blockDim.x) {
            const float4 v = s4[i];
            float4 r;
            float c;
            r.x = qdq_generic(a, v.x, Tc, qs, c); r.y = qdq_generic(a, v.y, Tc, qs, c);
            r.z = qdq_generic(a, v.z, Tc, qs, c); r.w = qdq_generic(a, v.w, Tc, qs, c);
            if (add) {
                const float4 old = y4[i];
                r.x = __fadd_rn(old.x, r.x); r.y = __fadd_rn(old.y, r.y); r.z = __fadd_rn(old.z, r.z); r.w = __fadd_rn(old.w, r.w);
            }
            y4[i] = r;
        }
    } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            float c;
            float r = qdq_generic(a, s_row[i], Tc, qs, c);
            if (add) r = __fadd_rn(yb[i], r);
            yb[i] = r;
        }
    }
}

// Eligible: weights viewed as (1, rows, inner) with short rows.  *done = 0 -> caller uses reduce + sweep.
template <bool IS_MAX>
[[maybe_unused]] static int launch_rows_fused(b2q_ctx* ctx, const float* x, float* y, int64_t rows, int64_t inner,
                                              Prescale ps, FoldBias fb, UpdateArgs u, QdqArgs a, int clip_with_fresh,
                                              cudaStream_t st, int* done) {
    *done = 0;
    if (rows < 2 || a.codes != nullptr || u.stat_out != nullptr) return 0;
    if (inner >= B2Q_ROWS_CTA_MIN_INNER && inner <= B2Q_ROWS_CTA_MAX_INNER && rows <= 0x7fffffff) {
        const size_t smem = (size_t)inner * sizeof(float);
        bool& optin_done = IS_MAX ? ctx->rows_cta_optin[0] : ctx->rows_cta_optin[1];   // per device (per ctx)
        if (smem > 48 * 1024 && !optin_done) {
            B2Q_CHECK_CUDA(cudaFuncSetAttribute(rows_cta_kernel<IS_MAX>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                B2Q_ROWS_CTA_MAX_INNER * (int)sizeof(float)));
            optin_done = true;
        }
        b2q_timed_launch tl(ctx, B2Q_KIND_OTHER, 8.0 * (double)(rows * inner), st);
        b2q_launch_smem(ctx, rows_cta_kernel<IS_MAX>, (unsigned)rows, B2Q_THREADS, smem, st, x, y, rows, inner, ps, fb, u, a,
                        clip_with_fresh);
        B2Q_LAUNCH_CHECK(ctx);
        *done = 1;
        return 0;
    }
    if (inner > B2Q_ROWS_FUSED_MAX_INNER) return 0;
    const int nw = B2Q_THREADS / 32;
    const unsigned grid = (unsigned)((rows + nw - 1) / nw);
    b2q_timed_launch tl(ctx, B2Q_KIND_OTHER, 8.0 * (double)(rows * inner), st);
    b2q_launch(ctx, rows_fused_kernel<IS_MAX>, (unsigned)grid, B2Q_THREADS, st, x, y, rows, inner, ps, fb, u, a, clip_with_fresh);
    B2Q_LAUNCH_CHECK(ctx);
    *done = 1;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// K5: straight-through backward (copy / accumulate), K6: masked backward.
// ------------------------------------------------------------------------------------------------
template <int MASK>
__device__ __forceinline__ float mask_grad(float x, float dy, float T) {
    // multiplications by 0.0/1.0 exactly as the reference does (keeps -0.0 and NaN propagation identical)
    if (MASK == B2Q_MASK_OPEN)
        return __fmul_rn(__fmul_rn(dy, (x > -T) ? 1.f : 0.f), (x < T) ? 1.f : 0.f);
    if (MASK == B2Q_MASK_ABS_LE) return __fmul_rn(dy, (fabsf(x) <= T) ? 1.f : 0.f);
    if (MASK == B2Q_MASK_LT) return __fmul_rn(dy, (x < T) ? 1.f : 0.f);
    return dy;
}

// MASK == 0: plain STE copy (x unused).
template <int MASK, bool ADD, int UNROLL, int LDPOL, int STPOL>
__global__ void __launch_bounds__(B2Q_THREADS)
bwd_flat_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, FlatSplit sp,
                const float* thr, float thr_imm) {
    b2q_pdl_sync();
    const float T = (MASK != 0) ? (thr ? __ldg(thr) : thr_imm) : 0.f;
    const float* xb = x + sp.head;
    const float* gb = dy + sp.head;
    float* ob = dx + sp.head;
    const int64_t tile = (int64_t)B2Q_THREADS * UNROLL;
    const int64_t ntiles = (sp.n8 + tile - 1) / tile;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t base = t * tile + threadIdx.x;
        f8 xv[UNROLL], gv[UNROLL], ov[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) {
            const int64_t i = base + (int64_t)k * B2Q_THREADS;
            if (i < sp.n8) {
                gv[k] = ld_f8<LDPOL>(gb + 8 * i);
                if (MASK != 0) xv[k] = ld_f8<LDPOL>(xb + 8 * i);
                if (ADD) ov[k] = ld_f8<0>(ob + 8 * i);
            }
        }
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) {
            const int64_t i = base + (int64_t)k * B2Q_THREADS;
            if (i < sp.n8) {
                f8 r = gv[k];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (MASK != 0) r.v[j] = mask_grad<MASK>(xv[k].v[j], gv[k].v[j], T);
                    if (ADD) r.v[j] = __fadd_rn(ov[k].v[j], r.v[j]);
                }
                st_f8<STPOL>(ob + 8 * i, r);
            }
        }
    }
    if (blockIdx.x == 0) {
        const int64_t tid = threadIdx.x;
        int64_t idx = -1;
        if (tid < sp.head) idx = tid;
        else if (tid - sp.head < sp.tail) idx = sp.head + 8 * sp.n8 + (tid - sp.head);
        if (idx >= 0) {
            float r = (MASK != 0) ? mask_grad<MASK>(x[idx], dy[idx], T) : dy[idx];
            if (ADD) r = __fadd_rn(dx[idx], r);
            dx[idx] = r;
        }
    }
}

template <int MASK, bool ADD>
__global__ void __launch_bounds__(128)
bwd_seg_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, SegPlan pl,
               const float* thr, float thr_imm) {
    b2q_pdl_sync();
    const SegPiece pc = seg_piece(pl);
    const float T = thr ? __ldg(thr + pc.g) : thr_imm;
    for (int64_t o = pc.o0; o < pc.o1; ++o) {
        const int64_t off = (o * pl.groups + pc.g) * pl.inner;
        if (pl.vec == 4) {
            const float4* x4 = reinterpret_cast<const float4*>(x + off);
            const float4* g4 = reinterpret_cast<const float4*>(dy + off);
            float4* o4 = reinterpret_cast<float4*>(dx + off);
            const int64_t end = pc.i1 >> 2;
            for (int64_t i0 = (pc.i0 >> 2) + threadIdx.x; i0 < end; i0 += 2 * (int64_t)blockDim.x) {
                float4 xv[2], gv[2];
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int64_t j = i0 + (int64_t)k * blockDim.x;
                    if (j < end) { gv[k] = g4[j]; if (MASK != 0) xv[k] = x4[j]; }
                }
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int64_t j = i0 + (int64_t)k * blockDim.x;
                    if (j >= end) break;
                    float4 r = gv[k];
                    if (MASK != 0) {
                        r.x = mask_grad<MASK>(xv[k].x, gv[k].x, T); r.y = mask_grad<MASK>(xv[k].y, gv[k].y, T);
                        r.z = mask_grad<MASK>(xv[k].z, gv[k].z, T); r.w = mask_grad<MASK>(xv[k].w, gv[k].w, T);
                    }
                    if (ADD) {
                        const float4 old = o4[j];
                        r.x = __fadd_rn(old.x, r.x); r.y = __fadd_rn(old.y, r.y); r.z = __fadd_rn(old.z, r.z); r.w = __fadd_rn(old.w, r.w);
                    }
                    o4[j] = r;
                }
            }
            continue;
        }
        for (int64_t i = pc.i0 + threadIdx.x; i < pc.i1; i += blockDim.x) {
            float r = (MASK != 0) ? mask_grad<MASK>(x[off + i], dy[off + i], T) : dy[off + i];
            if (ADD) r = __fadd_rn(dx[off + i], r);
            dx[off + i] = r;
        }
    }
}

// the same flattened walk for the masked / straight-through backward of grouped activations with short rows
template <int MASK, bool ADD, int V>
__global__ void __launch_bounds__(B2Q_THREADS)
bwd_seg_flatidx_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx, SegPlan pl,
                       const float* thr, float thr_imm) {
    b2q_pdl_sync();
    constexpr int U = (V == 4) ? 2 : 4;
    const SegPiece pc = seg_piece(pl);
    const float T = thr ? __ldg(thr + pc.g) : thr_imm;
    FlatWalk w = flat_walk(pl, pc, V);
    for (unsigned long long t = threadIdx.x; t < w.total; t += (unsigned long long)U * blockDim.x) {
        float xv[U][V], gv[U][V];
        int64_t off[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            off[k] = w.base + (int64_t)w.o * w.ostride + (int64_t)w.i * V;
            if (t + (unsigned long long)k * blockDim.x < w.total) {
                if (V == 4) {
                    const float4 g4 = *reinterpret_cast<const float4*>(dy + off[k]);
                    gv[k][0] = g4.x; gv[k][V > 1 ? 1 : 0] = g4.y; gv[k][V > 2 ? 2 : 0] = g4.z; gv[k][V > 3 ? 3 : 0] = g4.w;
                    if (MASK != 0) {
                        const float4 x4 = *reinterpret_cast<const float4*>(x + off[k]);
                        xv[k][0] = x4.x; xv[k][V > 1 ? 1 : 0] = x4.y; xv[k][V > 2 ? 2 : 0] = x4.z; xv[k][V > 3 ? 3 : 0] = x4.w;
                    }
                } else {
                    gv[k][0] = dy[off[k]];
                    if (MASK != 0) xv[k][0] = x[off[k]];
                }
            } else {
                off[k] = -1;
            }
            flat_walk_next(w);
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            if (off[k] < 0) continue;
            float r[V];
#pragma unroll
            for (int e = 0; e < V; ++e) {
                r[e] = (MASK != 0) ? mask_grad<MASK>(xv[k][e], gv[k][e], T) : gv[k][e];
                if (ADD) r[e] = __fadd_rn(dx[off[k] + e], r[e]);
            }
            if (V == 4) *reinterpret_cast<float4*>(dx + off[k]) = make_float4(r[0], r[V > 1 ? 1 : 0], r[V > 2 ? 2 : 0], r[V > 3 ? 3 : 0]);
            else dx[off[k]] = r[0];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
static inline bool same_misalignment(const void* a, const void* b) {
    return (((uintptr_t)a ^ (uintptr_t)b) & 31) == 0;
}

[[maybe_unused]] static int launch_qdq(b2q_ctx* ctx, const float* x, float* y, int64_t outer, int64_t groups, int64_t inner,
                      Prescale ps, FoldBias fb, QdqArgs a, cudaStream_t st) {
    B2Q_REQUIRE(outer >= 1 && groups >= 1 && inner >= 1, "empty tensor");
    B2Q_REQUIRE(a.req == B2Q_REQ_WRITE || a.req == B2Q_REQ_INPLACE || a.req == B2Q_REQ_ADD, "bad req");
    B2Q_REQUIRE(a.clip_mode >= B2Q_CLIP_NONE && a.clip_mode <= B2Q_CLIP_WHERE_LT, "unknown clip_mode");
    const int64_t n = outer * groups * inner;
    if (groups == 1 && ps.gamma == nullptr) {
        FlatSplit sp = b2q_flat_split(x, n);
        bool ok = same_misalignment(x, y) && (!a.codes || ((((uintptr_t)x >> 2) & 7) == (((uintptr_t)a.codes >> 2) & 7)));
        if (ok && sp.head <= B2Q_THREADS) {
            const bool hot = a.do_round && a.req != B2Q_REQ_ADD && !a.codes;
            b2q_timed_launch tl(ctx, hot ? B2Q_KIND_QDQ_HOT : B2Q_KIND_OTHER, 8.0 * (double)n, st);
            if (hot) {
                const int64_t grid = b2q_flat_grid(ctx, sp.n8, B2Q_QDQ_UNROLL);
                DeferredUpdate none = {};
                const bool stream_out = n * 4 > B2Q_STREAM_BYTES;
                const int rev = (ctx->reverse && n * 4 > ((long long)ctx->reverse_min_mb << 20)) ? 1 : 0;
#define B2Q_HOT(C, S) b2q_launch(ctx, qdq_flat_hot_kernel<C, B2Q_QDQ_UNROLL, B2Q_QDQ_LDPOL, S, false>, \
                                 (unsigned)grid, B2Q_THREADS, st, x, y, sp, a, rev, none, 0)
#define B2Q_HOT_S(C) do { if (stream_out) B2Q_HOT(C, 1); else B2Q_HOT(C, B2Q_QDQ_STPOL); } while (0)
                switch (a.clip_mode) {
                    case B2Q_CLIP_SYM: B2Q_HOT_S(B2Q_CLIP_SYM); break;
                    case B2Q_CLIP_WHERE_LE: B2Q_HOT_S(B2Q_CLIP_WHERE_LE); break;
                    case B2Q_CLIP_ZERO_T: B2Q_HOT_S(B2Q_CLIP_ZERO_T); break;
                    case B2Q_CLIP_PACT: B2Q_HOT_S(B2Q_CLIP_PACT); break;
                    case B2Q_CLIP_WHERE_LT: B2Q_HOT_S(B2Q_CLIP_WHERE_LT); break;
                    default: B2Q_HOT_S(B2Q_CLIP_NONE); break;
                }
#undef B2Q_HOT_S
#undef B2Q_HOT
            } else {
                const int64_t grid = b2q_flat_grid(ctx, sp.n8, 2);
                b2q_launch(ctx, qdq_flat_generic_kernel<2>, (unsigned)grid, B2Q_THREADS, st, x, y, sp, a);
            }
            B2Q_LAUNCH_CHECK(ctx);
            return 0;
        }
        outer = 1; inner = n;  // mutually misaligned buffers: scalar segmented path
    }
    SegPlan pl = b2q_seg_plan(x, y, outer, groups, inner, ctx->num_sms * 16);
    const unsigned grid = (unsigned)(groups * pl.S * pl.P);
    b2q_timed_launch tl(ctx, B2Q_KIND_OTHER, 8.0 * (double)n, st);
    // large grouped tensors: hot inner loop per piece (needs 32-byte aligned ranges: inner and part multiples of 8)
    if (ps.gamma == nullptr && fb.bias == nullptr && a.thr != nullptr && a.do_round && a.req != B2Q_REQ_ADD && !a.codes &&
        (inner % 8 == 0) && (pl.part % 8 == 0) && ((((uintptr_t)x) | ((uintptr_t)y)) & 31) == 0 && n >= (1 << 20)) {
        switch (a.clip_mode) {
            case B2Q_CLIP_WHERE_LE: b2q_launch(ctx, qdq_seg_hot_kernel<B2Q_CLIP_WHERE_LE>, grid, B2Q_THREADS, st, x, y, pl, a); break;
            case B2Q_CLIP_SYM: b2q_launch(ctx, qdq_seg_hot_kernel<B2Q_CLIP_SYM>, grid, B2Q_THREADS, st, x, y, pl, a); break;
            default: b2q_launch(ctx, qdq_seg_hot_kernel<B2Q_CLIP_NONE>, grid, B2Q_THREADS, st, x, y, pl, a); break;
        }
        if (a.clip_mode == B2Q_CLIP_WHERE_LE || a.clip_mode == B2Q_CLIP_SYM || a.clip_mode == B2Q_CLIP_NONE) {
            B2Q_LAUNCH_CHECK(ctx);
            return 0;
        }
    }
    // large grouped activations with short rows (14x14 / 7x7 maps): flattened (row, element) walk
    if (ps.gamma == nullptr && fb.bias == nullptr && !a.codes && pl.P == 1 && inner < 2048 && n >= (1 << 18)) {
        if (pl.vec == 4) b2q_launch(ctx, qdq_seg_flatidx_kernel<4>, grid, B2Q_THREADS, st, x, y, pl, a);
        else b2q_launch(ctx, qdq_seg_flatidx_kernel<1>, grid, B2Q_THREADS, st, x, y, pl, a);
        B2Q_LAUNCH_CHECK(ctx);
        return 0;
    }
    if (pl.vec == 4) b2q_launch(ctx, qdq_seg_kernel<4>, grid, 128, st, x, y, pl, ps, fb, a);
    else b2q_launch(ctx, qdq_seg_kernel<1>, grid, 128, st, x, y, pl, ps, fb, a);
    B2Q_LAUNCH_CHECK(ctx);
    return 0;
}

// Fused whole-tensor forward: reduction (partials only) + QDQ sweep that finishes the threshold update itself.
// Eligible when the sweep is the hot kernel (clip none / symmetric, rounding, req write) and the buffers allow the
// 256-bit path; *done = 0 tells the caller to take the generic reduce -> update -> sweep route instead.
template <bool IS_MAX>
[[maybe_unused]] static int launch_fused_flat_fwd(b2q_ctx* ctx, b2q_slot* slot, const float* x, float* y, int64_t n,
                                                  UpdateArgs u, float qlevel, int clip_mode, int clip_with_fresh,
                                                  cudaStream_t st, int* done) {
    *done = 0;
    // Sums keep the last-block finalisation: a deferred sum would make every one-tile block of the sweep re-combine
    // hundreds of double partials; the max needs only one tagged word.
    if (!IS_MAX) return 0;
    if (!ctx->deferred || !(clip_mode == B2Q_CLIP_NONE || clip_mode == B2Q_CLIP_SYM)) return 0;
    if (u.stat_out != nullptr) return 0;
    FlatSplit sp = b2q_flat_split(x, n);
    if (!same_misalignment(x, y) || sp.head > B2Q_THREADS) return 0;
    int np = 0;
    int rc = launch_reduce_deferred<IS_MAX>(ctx, slot, x, n, u, st, &np);
    if (rc || np == 0) return rc;
    DeferredUpdate d;
    d.partial = slot->partial;
    d.max64 = &slot->max64;
    d.epoch = &slot->epoch;
    d.aux_old = slot->scale;
    d.n_partials = np;
    d.is_max = IS_MAX ? 1 : 0;
    d.count = (float)n;
    d.u = u;
    QdqArgs a = {nullptr, nullptr, 0.f, 0.f, qlevel, ctx->fast_div, nullptr, clip_mode, 1, B2Q_REQ_WRITE};
    const int64_t grid = b2q_flat_grid(ctx, sp.n8, B2Q_QDQ_UNROLL);
    {
        b2q_timed_launch tl(ctx, B2Q_KIND_QDQ_HOT, 8.0 * (double)n, st);
        const bool stream_out = n * 4 > B2Q_STREAM_BYTES;
        const int rev = (ctx->reverse && n * 4 > ((long long)ctx->reverse_min_mb << 20)) ? 1 : 0;
#define B2Q_HOT(C, S) b2q_launch(ctx, qdq_flat_hot_kernel<C, B2Q_QDQ_UNROLL, B2Q_QDQ_LDPOL, S, true>, \
                                 (unsigned)grid, B2Q_THREADS, st, x, y, sp, a, rev, d, clip_with_fresh)
        if (clip_mode == B2Q_CLIP_SYM) { if (stream_out) B2Q_HOT(B2Q_CLIP_SYM, 1); else B2Q_HOT(B2Q_CLIP_SYM, B2Q_QDQ_STPOL); }
        else { if (stream_out) B2Q_HOT(B2Q_CLIP_NONE, 1); else B2Q_HOT(B2Q_CLIP_NONE, B2Q_QDQ_STPOL); }
#undef B2Q_HOT
        B2Q_LAUNCH_CHECK(ctx);
    }
    *done = 1;
    return 0;
}

template <int MASK>
static int launch_bwd_mask(b2q_ctx* ctx, const float* x, const float* dy, float* dx, int64_t outer, int64_t groups,
                           int64_t inner, const float* thr, float thr_imm, int req, cudaStream_t st) {
    const int64_t n = outer * groups * inner;
    const bool add = (req == B2Q_REQ_ADD);
    if (groups == 1) {
        FlatSplit sp = b2q_flat_split(dy, n);
        bool ok = same_misalignment(dy, dx) && (MASK == 0 || same_misalignment(dy, x));
        if (ok && sp.head <= B2Q_THREADS) {
            const int64_t grid = b2q_flat_grid(ctx, sp.n8, B2Q_BWD_UNROLL);
            b2q_timed_launch tl(ctx, MASK == 0 ? B2Q_KIND_BWD_STE : B2Q_KIND_BWD_MASK, (MASK == 0 ? 8.0 : 12.0) * (double)n, st);
            const bool stream_out = n * 4 > B2Q_STREAM_BYTES;
            if (add) b2q_launch(ctx, bwd_flat_kernel<MASK, true, B2Q_BWD_UNROLL, B2Q_BWD_LDPOL, B2Q_BWD_STPOL>,
                                (unsigned)grid, B2Q_THREADS, st, x, dy, dx, sp, thr, thr_imm);
            else if (stream_out) b2q_launch(ctx, bwd_flat_kernel<MASK, false, B2Q_BWD_UNROLL, B2Q_BWD_LDPOL, 1>,
                                            (unsigned)grid, B2Q_THREADS, st, x, dy, dx, sp, thr, thr_imm);
            else b2q_launch(ctx, bwd_flat_kernel<MASK, false, B2Q_BWD_UNROLL, B2Q_BWD_LDPOL, B2Q_BWD_STPOL>,
                            (unsigned)grid, B2Q_THREADS, st, x, dy, dx, sp, thr, thr_imm);
            B2Q_LAUNCH_CHECK(ctx);
            return 0;
        }
        outer = 1; inner = n;
    }
    B2Q_REQUIRE(thr != nullptr || groups == 1, "grouped mask needs a device threshold vector");
    SegPlan pl = b2q_seg_plan(dy, dx, outer, groups, inner, ctx->num_sms * 16);
    if (MASK != 0 && (((uintptr_t)x) & 15)) pl.vec = 1;
    const unsigned grid = (unsigned)(groups * pl.S * pl.P);
    if (pl.P == 1 && inner < 2048 && n >= (1 << 18)) {   // short rows: flattened (row, element) walk
        if (pl.vec == 4) {
            if (add) b2q_launch(ctx, bwd_seg_flatidx_kernel<MASK, true, 4>, grid, B2Q_THREADS, st, x, dy, dx, pl, thr, thr_imm);
            else b2q_launch(ctx, bwd_seg_flatidx_kernel<MASK, false, 4>, grid, B2Q_THREADS, st, x, dy, dx, pl, thr, thr_imm);
        } else {
            if (add) b2q_launch(ctx, bwd_seg_flatidx_kernel<MASK, true, 1>, grid, B2Q_THREADS, st, x, dy, dx, pl, thr, thr_imm);
            else b2q_launch(ctx, bwd_seg_flatidx_kernel<MASK, false, 1>, grid, B2Q_THREADS, st, x, dy, dx, pl, thr, thr_imm);
        }
        B2Q_LAUNCH_CHECK(ctx);
        return 0;
    }
    if (add) b2q_launch(ctx, bwd_seg_kernel<MASK, true>, grid, 128, st, x, dy, dx, pl, thr, thr_imm);
    else b2q_launch(ctx, bwd_seg_kernel<MASK, false>, grid, 128, st, x, dy, dx, pl, thr, thr_imm);
    B2Q_LAUNCH_CHECK(ctx);
    return 0;
}
