// K1/K2/K7: max|x| and sum|x| reductions with the threshold update fused into the last block.
#pragma once
#include "b2q_common.cuh"

// Compile-time tuning of the flat (whole-tensor) kernels; chosen from tools/sweep.cu runs on B200
// (profiles/).  UNROLL = independent 256-bit accesses in flight per thread.
#ifndef B2Q_REDUCE_UNROLL
#define B2Q_REDUCE_UNROLL 4
#endif
#ifndef B2Q_REDUCE_LDPOL
#define B2Q_REDUCE_LDPOL 0
#endif
#ifndef B2Q_REDUCE_BPS
#define B2Q_REDUCE_BPS 4     // blocks per SM for the flat reduction (sweep: 4 is best from 50 MB up)
#endif

// A flat float32 array split for 256-bit access: `head` scalars until 32-byte alignment, n8 groups of eight
// floats, `tail` scalars.
struct FlatSplit {
    int64_t head, n8, tail;
};

static inline FlatSplit b2q_flat_split(const void* p, int64_t n) {
    FlatSplit s;
    int64_t mis = (int64_t)(((uintptr_t)p >> 2) & 7);
    s.head = mis ? (8 - mis) : 0;
    if (((uintptr_t)p & 3) != 0) s.head = n;  // not float aligned: all scalar
    if (s.head > n) s.head = n;
    s.n8 = (n - s.head) / 8;
    s.tail = n - s.head - 8 * s.n8;
    return s;
}

template <bool IS_MAX>
__device__ __forceinline__ void acc1(double& a, float& m, float v) {
    if (IS_MAX) m = fmax_nan(m, fabsf(v)); else a += (double)fabsf(v);
}

// ICVT selects where the float -> double conversions of a sum run: false = the conversion instruction (XU pipe; best
// for the register-staged loops, where occupancy is the limit), true = the integer pipe (f2d_abs_scaled,
// b2q_common.cuh; for the TMA-staged loops, where the XU pipe would be).  Same bits either way (b2q_selftest(5)).
template <bool IS_MAX, bool ICVT = false>
__device__ __forceinline__ void acc8(double& a, float& m, const f8& r) {
    if (IS_MAX) {
        const float m0 = fmax_nan(fabsf(r.v[0]), fabsf(r.v[1])), m1 = fmax_nan(fabsf(r.v[2]), fabsf(r.v[3]));
        const float m2 = fmax_nan(fabsf(r.v[4]), fabsf(r.v[5])), m3 = fmax_nan(fabsf(r.v[6]), fabsf(r.v[7]));
        m = fmax_nan(m, fmax_nan(fmax_nan(m0, m1), fmax_nan(m2, m3)));
    } else if (!ICVT) {
        // double accumulation: every float32 -> double conversion and each pair sum below 2^53 ulps is exact
        const double s0 = (double)fabsf(r.v[0]) + (double)fabsf(r.v[1]);
        const double s1 = (double)fabsf(r.v[2]) + (double)fabsf(r.v[3]);
        const double s2 = (double)fabsf(r.v[4]) + (double)fabsf(r.v[5]);
        const double s3 = (double)fabsf(r.v[6]) + (double)fabsf(r.v[7]);
        a += (s0 + s1) + (s2 + s3);
    } else {
        // the same tree at scale 2^-896 -- which rounds exactly as the unscaled tree does, every operand being a
        // multiple of 2^-149 * 2^-896 -- rescaled once per word
        unsigned t = 0;   // bit 31 set iff some exponent field is 255
#pragma unroll
        for (int j = 0; j < 8; ++j) t |= (__float_as_uint(r.v[j]) & 0x7fffffffu) + 0x00800000u;
        if (t & 0x80000000u) {   // Inf / NaN in this word: the real conversion (they propagate as in the reference's sum)
            const double s0 = f2d_cvt(fabsf(r.v[0])) + f2d_cvt(fabsf(r.v[1])), s1 = f2d_cvt(fabsf(r.v[2])) + f2d_cvt(fabsf(r.v[3]));
            const double s2 = f2d_cvt(fabsf(r.v[4])) + f2d_cvt(fabsf(r.v[5])), s3 = f2d_cvt(fabsf(r.v[6])) + f2d_cvt(fabsf(r.v[7]));
            a += (s0 + s1) + (s2 + s3);
        } else {
            bool unused = false;
            const double s0 = f2d_abs_scaled(r.v[0], unused) + f2d_abs_scaled(r.v[1], unused);
            const double s1 = f2d_abs_scaled(r.v[2], unused) + f2d_abs_scaled(r.v[3], unused);
            const double s2 = f2d_abs_scaled(r.v[4], unused) + f2d_abs_scaled(r.v[5], unused);
            const double s3 = f2d_abs_scaled(r.v[6], unused) + f2d_abs_scaled(r.v[7], unused);
            a += ((s0 + s1) + (s2 + s3)) * b2q_two_p896();
        }
    }
}

// sum and sum of squares of one 256-bit word (BatchNorm_v1 statistics, b2q_bnfold.cu): one conversion, one DADD (pair
// tree) and one DFMA (two independent chains) per element
template <bool ICVT>
__device__ __forceinline__ void acc8_sq(double& s, double& q0, double& q1, const f8& r) {
    double d[8];
    if (!ICVT) {
#pragma unroll
        for (int e = 0; e < 8; ++e) d[e] = (double)r.v[e];
    } else {
        unsigned t = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) t |= (__float_as_uint(r.v[e]) & 0x7fffffffu) + 0x00800000u;
        if (t & 0x80000000u) {
#pragma unroll
            for (int e = 0; e < 8; ++e) d[e] = f2d_cvt(r.v[e]);
        } else {
            bool unused = false;
#pragma unroll
            for (int e = 0; e < 8; ++e) d[e] = f2d_scaled(r.v[e], unused) * b2q_two_p896();   // exact
        }
    }
    s += ((d[0] + d[1]) + (d[2] + d[3])) + ((d[4] + d[5]) + (d[6] + d[7]));
    q0 = fma(d[0], d[0], q0); q1 = fma(d[1], d[1], q1);
    q0 = fma(d[2], d[2], q0); q1 = fma(d[3], d[3], q1);
    q0 = fma(d[4], d[4], q0); q1 = fma(d[5], d[5], q1);
    q0 = fma(d[6], d[6], q0); q1 = fma(d[7], d[7], q1);
}

// The same sums with the smallest register footprint: each element is converted, added and squared before the next
// one is touched (two alternating chains per sum), so only the loaded words and four accumulators stay live and more
// loads fit in flight per SM -- what the register-staged segmented loops are short of (profiles/r02i_bnstat_pipes.md).
__device__ __forceinline__ void acc1_sq_asm(double& s, double& q, float x) {
    // one statement, so that the compiler cannot convert a whole batch of words first and keep the doubles alive
    asm volatile("{\n\t.reg .f64 d;\n\tcvt.f64.f32 d, %2;\n\tadd.rn.f64 %0, %0, d;\n\tfma.rn.f64 %1, d, d, %1;\n\t}"
                 : "+d"(s), "+d"(q)
                 : "f"(x));
}

// the same step with the conversion on the integer pipe (f2d_scaled, b2q_common.cuh); finite x only
__device__ __forceinline__ void acc1_sq_asm_icvt(double& s, double& q, float x) {
    asm volatile(
        "{\n\t.reg .f64 d;\n\t.reg .b32 a, hi, lo;\n\t"
        "and.b32 a, %2, 0x7fffffff;\n\tshr.u32 hi, a, 3;\n\tlop3.b32 hi, hi, %2, 0x80000000, 0xf8;\n\tshl.b32 lo, %2, 29;\n\t"
        "mov.b64 d, {lo, hi};\n\tmul.rn.f64 d, d, %3;\n\tadd.rn.f64 %0, %0, d;\n\tfma.rn.f64 %1, d, d, %1;\n\t}"
        : "+d"(s), "+d"(q)
        : "r"(__float_as_uint(x)), "d"(b2q_two_p896()));
}

// CVT: 0 = every conversion on the XU pipe, 1 = every conversion on the integer pipe, 2 = half and half (elements 0-3 /
// 4-7 of the word), which loads both pipes
template <int CVT>
__device__ __forceinline__ void acc8_sq_seq(double& s0, double& s1, double& q0, double& q1, const f8& r) {
    const int first_int = CVT == 1 ? 0 : (CVT == 2 ? 4 : 8);   // elements [first_int, 8) go through the integer pipe
    bool fast = true;
    if (CVT != 0) {
        unsigned t = 0;   // bit 31 set iff some exponent field is 255 (Inf / NaN): those words take the real conversion
#pragma unroll
        for (int e = first_int; e < 8; ++e) t |= (__float_as_uint(r.v[e]) & 0x7fffffffu) + 0x00800000u;
        fast = (t & 0x80000000u) == 0;
    }
#pragma unroll
    for (int e = 0; e < first_int; e += 2) {
        acc1_sq_asm(s0, q0, r.v[e]);
        acc1_sq_asm(s1, q1, r.v[e + 1]);
    }
    if (CVT != 0) {
        if (fast) {
#pragma unroll
            for (int e = first_int; e < 8; e += 2) {
                acc1_sq_asm_icvt(s0, q0, r.v[e]);
                acc1_sq_asm_icvt(s1, q1, r.v[e + 1]);
            }
        } else {
#pragma unroll
            for (int e = first_int; e < 8; e += 2) {
                acc1_sq_asm(s0, q0, r.v[e]);
                acc1_sq_asm(s1, q1, r.v[e + 1]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Whole-tensor reduction (groups == 1): the hot kernel for activations.
// Grid-stride over tiles of THREADS*UNROLL 256-bit words in ASCENDING address order, so that the most
// recently read part of x is what remains in L2 for the QDQ sweep that follows (which walks descending).
// ------------------------------------------------------------------------------------------------
template <bool IS_MAX, int UNROLL, int LDPOL, bool FINALIZE>
__global__ void __launch_bounds__(B2Q_THREADS)
reduce_flat_kernel(const float* __restrict__ x, FlatSplit sp, b2q_slot* slot, UpdateArgs u, float count) {
    b2q_pdl_sync();
    __shared__ double smem[32];
    __shared__ unsigned int s_ticket;
    double acc = 0.0;
    float mx = 0.f;
    const float* xb = x + sp.head;
    const int64_t tile = (int64_t)B2Q_THREADS * UNROLL;
    const int64_t ntiles = (sp.n8 + tile - 1) / tile;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t base = t * tile + threadIdx.x;
        f8 v[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) {
            const int64_t i = base + (int64_t)k * B2Q_THREADS;
            if (i < sp.n8) v[k] = ld_f8<LDPOL>(xb + 8 * i);
            else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[k].v[j] = 0.f;
            }
        }
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) acc8<IS_MAX>(acc, mx, v[k]);
    }
    if (blockIdx.x == 0) {  // the (at most 14) unaligned scalars
        if ((int64_t)threadIdx.x < sp.head) acc1<IS_MAX>(acc, mx, x[threadIdx.x]);
        if ((int64_t)threadIdx.x < sp.tail) acc1<IS_MAX>(acc, mx, x[sp.head + 8 * sp.n8 + threadIdx.x]);
    }
    double r = block_reduce<IS_MAX>(IS_MAX ? (double)mx : acc, smem);
    if (!FINALIZE) {
        // deferred update: the consumer kernel combines the partials (see DeferredUpdate)
        if (threadIdx.x == 0) {
            if (IS_MAX) {
                // tag = slot->epoch + 1: strictly newer than anything the word holds, so no reset is ever needed
                const unsigned long long tag = (unsigned long long)(slot->epoch + 1u) << 32;
                atomicMax(&slot->max64, tag | __float_as_uint((float)r));
            } else {
                slot->partial[blockIdx.x] = r;
            }
            if (blockIdx.x == 0 && u.aux) slot->scale[0] = u.aux[0];   // snapshot of the old threshold
            // peer exchange: this call's sequence number (one thread of the whole grid; kernels of a stream are ordered)
            if (blockIdx.x == 0 && u.seq_counter) *u.seq_counter = *u.seq_counter + 1u;
        }
        return;
    }
    if (gridDim.x == 1) {   // single block: nothing to combine, no fence / ticket round trips
        if (threadIdx.x == 0) apply_update(u, 0, IS_MAX ? (float)r : __fdiv_rn((float)r, count));
        return;
    }
    if (threadIdx.x == 0) {
        slot->partial[blockIdx.x] = r;
        s_ticket = b2q_take_ticket(&slot->ticket, gridDim.x - 1);
    }
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    // ---- last block: combine in fixed order, update the threshold ----
    double a = 0.0;
    float m = 0.f;
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
        double p = __ldcg(&slot->partial[i]);
        if (IS_MAX) m = fmax_nan(m, (float)p); else a += p;
    }
    double tot = block_reduce<IS_MAX>(IS_MAX ? (double)m : a, smem);
    if (threadIdx.x == 0) {
        float stat = IS_MAX ? (float)tot : __fdiv_rn((float)tot, count);
        apply_update(u, 0, stat);
        slot->ticket = 0;
    }
}

// ------------------------------------------------------------------------------------------------
// Segmented reduction over a (outer, groups, inner) view.  Piece (g, s, p) covers outer-range s and
// inner-range p of group g; partial[(g*S + s)*P + p].  Optional per-row prescale (fold-BN).
// ------------------------------------------------------------------------------------------------
struct SegPlan {
    int64_t outer, groups, inner;
    int S, P;            // splits of outer / inner
    int64_t part;        // inner elements per p-part (multiple of 8)
    int vec;             // 1 or 4
};

static inline SegPlan b2q_seg_plan(const void* x, const void* y, int64_t outer, int64_t groups, int64_t inner,
                                   int target_pieces) {
    SegPlan pl;
    pl.outer = outer; pl.groups = groups; pl.inner = inner;
    bool aligned = (((uintptr_t)x & 15) == 0) && (((uintptr_t)y & 15) == 0) && (inner % 4 == 0);
    pl.vec = aligned ? 4 : 1;
    int64_t want = (target_pieces + groups - 1) / groups;  // pieces per group
    if (want < 1) want = 1;
    int64_t S = outer < want ? outer : want;
    if (S < 1) S = 1;
    int64_t wantP = (want + S - 1) / S;
    int64_t min_part = 4096;  // do not cut rows into pieces smaller than this many elements
    int64_t maxP = (inner + min_part - 1) / min_part;
    int64_t P = wantP < maxP ? wantP : maxP;
    if (P < 1) P = 1;
    while (groups * S * P > B2Q_MAX_PIECES && P > 1) --P;
    while (groups * S * P > B2Q_MAX_PIECES && S > 1) --S;
    int64_t part = (inner + P - 1) / P;
    part = (part + 7) & ~(int64_t)7;   // 32-byte granules: pieces stay 256-bit aligned
    P = (inner + part - 1) / part;
    if (P < 1) P = 1;
    pl.S = (int)S; pl.P = (int)P; pl.part = part;
    return pl;
}

struct SegPiece {
    int64_t g, o0, o1, i0, i1;
    int p;
};

__device__ __forceinline__ SegPiece seg_piece(const SegPlan& pl) {
    SegPiece sp;
    const int64_t piece = blockIdx.x;
    const int SP = pl.S * pl.P;
    sp.g = piece / SP;
    const int s = (int)((piece / pl.P) % pl.S);
    sp.p = (int)(piece % pl.P);
    sp.o0 = (pl.outer * s) / pl.S;
    sp.o1 = (pl.outer * (s + 1)) / pl.S;
    sp.i0 = (int64_t)sp.p * pl.part;
    sp.i1 = sp.i0 + pl.part;
    if (sp.i1 > pl.inner) sp.i1 = pl.inner;
    return sp;
}

// ------------------------------------------------------------------------------------------------
// TMA-staged accumulation of one piece of an (outer, groups, inner) view (option stream_reduce = 1; OFF by default).
// The register-staged loops below keep 2-4 256-bit loads per thread in flight: 60-130 KB per SM, and ncu shows the
// warps waiting on the scoreboard for L1TEX 56 % of the time at 49 % of DRAM throughput.  Here the piece streams through
// a ring of 16 KB shared-memory stages filled by bulk asynchronous copies (cp.async.bulk + mbarrier transaction counts,
// L2 evict-first hint): the data in flight (stages x 16 KB per block) costs no registers, and the threads only ever
// read shared memory.  A row of the piece (contiguous, a multiple of 32 bytes) is cut into sub-rows of at most one
// stage; short rows are packed several to a stage, one copy each, issued by the lanes of warp 0.  Because this is a
// reduction, which thread takes which 16 bytes of a stage is free: thread t reads the float4s t and t + 256,
// conflict-free.  MEASURED (profiles/r02i_bnstat_pipes.md): bit-identical results, but 10-25 % SLOWER than the
// register-staged loops at every ring depth -- a 16 KB stage is one loop iteration per thread, so the wait / barrier /
// refill round per stage costs more than the deeper prefetch gains, and deeper rings cost resident blocks.  Kept as a
// tested option, not as the default.
//   MODE 0: sum |x|   MODE 1: max |x|   MODE 2: sum x and sum x^2 (acc = sum, acc2 = sum of squares)
// ------------------------------------------------------------------------------------------------
#define B2Q_STREAM_CH 4096   // floats per stage (16 KB)

template <int MODE, bool ICVT>
__device__ __forceinline__ void seg_stream_accumulate(const float* __restrict__ x, const SegPlan& pl, const SegPiece& pc,
                                                      float* s_buf, unsigned long long* s_bar, int nst, double& acc,
                                                      double& acc2, float& mx) {
    const int64_t len = pc.i1 - pc.i0;                    // floats per row inside this piece (multiple of 8)
    const int64_t rows = pc.o1 - pc.o0;
    const float* base = x + (pc.o0 * pl.groups + pc.g) * pl.inner + pc.i0;
    const int64_t ostride = pl.groups * pl.inner;
    const int cpr = (int)((len + B2Q_STREAM_CH - 1) / B2Q_STREAM_CH);   // sub-rows per row
    const int L = cpr > 1 ? B2Q_STREAM_CH : (int)len;     // floats per (full) sub-row
    const int spf = cpr > 1 ? 1 : B2Q_STREAM_CH / L;      // sub-rows per fill
    const int64_t tsub = rows * cpr;
    const int64_t fills = (tsub + spf - 1) / spf;
    const int lane = threadIdx.x & 31;
    auto fill_floats = [&](int64_t f) -> int {
        const int64_t j0 = f * spf;
        if (cpr > 1) {
            const int64_t c = j0 % cpr;
            const int64_t n = len - c * B2Q_STREAM_CH;
            return (int)(n < B2Q_STREAM_CH ? n : B2Q_STREAM_CH);
        }
        const int64_t cnt = tsub - j0 < spf ? tsub - j0 : spf;
        return (int)cnt * L;
    };
    auto issue = [&](int64_t f) {   // all lanes of warp 0
        const int stage = (int)(f % nst);
        const int64_t j0 = f * spf;
        const int cnt = (int)(tsub - j0 < spf ? tsub - j0 : spf);
        const unsigned long long pol = l2_policy_evict_first();
        if (lane == 0) mbar_expect_tx(&s_bar[stage], (unsigned)fill_floats(f) * 4u);
        __syncwarp();
        for (int q = lane; q < cnt; q += 32) {
            const int64_t j = j0 + q;
            const int64_t row = j / cpr, c = j - row * cpr;
            int64_t n = cpr > 1 ? len - c * B2Q_STREAM_CH : (int64_t)L;
            if (n > B2Q_STREAM_CH) n = B2Q_STREAM_CH;
            bulk_g2s_hint(s_buf + (size_t)stage * B2Q_STREAM_CH + (size_t)q * L, base + row * ostride + c * B2Q_STREAM_CH,
                          (unsigned)n * 4u, &s_bar[stage], pol);
        }
    };
    if (threadIdx.x == 0) {
        for (int i = 0; i < nst; ++i) mbar_init(&s_bar[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        for (int64_t f = 0; f < fills && f < nst; ++f) issue(f);
    }
    double q0 = 0.0, q1 = 0.0;
    for (int64_t f = 0; f < fills; ++f) {
        const int stage = (int)(f % nst);
        mbar_wait(&s_bar[stage], (unsigned)((f / nst) & 1));
        const int nq = fill_floats(f) >> 2;               // float4s in this stage (even)
        const float4* s4 = reinterpret_cast<const float4*>(s_buf + (size_t)stage * B2Q_STREAM_CH);
        for (int q = threadIdx.x; q < nq; q += 2 * B2Q_THREADS) {
            const float4 a = s4[q];
            const float4 b = (q + B2Q_THREADS < nq) ? s4[q + B2Q_THREADS] : make_float4(0.f, 0.f, 0.f, 0.f);
            f8 v;
            v.v[0] = a.x; v.v[1] = a.y; v.v[2] = a.z; v.v[3] = a.w; v.v[4] = b.x; v.v[5] = b.y; v.v[6] = b.z; v.v[7] = b.w;
            if (MODE == 2) acc8_sq<ICVT>(acc, q0, q1, v);
            else acc8<MODE == 1, ICVT>(acc, mx, v);
        }
        __syncthreads();                                   // everyone is done with this stage: refill it
        if (threadIdx.x < 32 && f + nst < fills) issue(f + nst);
    }
    acc2 = q0 + q1;
}

// Rows that are NOT a multiple of eight floats (7x7 = 49, 14x14 = 196, ...) with 256-bit loads all the same: every
// ALIGNED 256-bit word that overlaps a row of the piece is loaded whole and the elements outside the row are zeroed
// (neutral for sums, sums of squares and max|x|; a select, so a neighbour's NaN does not leak in).  The (row, word) space is
// flattened into one index walked incrementally; U words in flight per thread.  x must be 32-byte aligned and hold
// n_total elements (the last word of the tensor is read element-wise if it would cross the end).  A 49-float row costs
// 7-8 words for 6.1 words of payload, and its edge words are shared with the neighbouring rows (another block's, an L2
// hit) -- against eight times fewer load instructions and eight times the bytes in flight of the scalar walk.
// Measured (profiles/r02v_bnstat_small_maps.md): batch statistics on 7x7 maps 0.25 -> 0.32 / 0.21 -> 0.25 of the copy
// peak; on 14x14 maps (rows a multiple of four floats) the 128-bit walk is faster (0.56 vs 0.49), and the grouped
// mean|x| / max|x| reductions do not gain -- so only the statistics kernel uses it, and only for rows that are not a
// multiple of four floats.
template <int U, typename F>
__device__ __forceinline__ void seg_masked_words(const float* __restrict__ x, int64_t n_total, const SegPlan& pl,
                                                 const SegPiece& pc, F&& f) {
    const int64_t len = pc.i1 - pc.i0;
    const unsigned wpr = (unsigned)((len + 14) >> 3);             // upper bound of the words overlapping one row
    const unsigned long long total = (unsigned long long)(pc.o1 - pc.o0) * wpr;
    const int64_t row0 = (pc.o0 * pl.groups + pc.g) * pl.inner + pc.i0;
    const int64_t ostride = pl.groups * pl.inner;
    const unsigned qT = blockDim.x / wpr, rT = blockDim.x % wpr;
    unsigned o = threadIdx.x / wpr, j = threadIdx.x % wpr;
    for (unsigned long long t = threadIdx.x; t < total; t += (unsigned long long)U * blockDim.x) {
        f8 v[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int64_t start = row0 + (int64_t)o * ostride;
            const int64_t wbase = (start & ~(int64_t)7) + 8 * (int64_t)j;
            const int lo = (int)(start - wbase);                   // elements [lo, hi) of the word belong to the row
            const int64_t hi64 = start + len - wbase;
            const int hi = hi64 > 8 ? 8 : (int)hi64;
            if (t + (unsigned long long)k * blockDim.x < total && hi > 0) {
                if (wbase + 8 <= n_total) {
                    v[k] = ld_f8<1>(x + wbase);
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[k].v[e] = (wbase + e < n_total) ? x[wbase + e] : 0.f;
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) v[k].v[e] = (e < lo || e >= hi) ? 0.f : v[k].v[e];
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) v[k].v[e] = 0.f;
            }
            j += rT; o += qT;
            if (j >= wpr) { j -= wpr; ++o; }
        }
#pragma unroll
        for (int k = 0; k < U; ++k) f(v[k]);
    }
}

struct Prescale {  // fold-BN factor gamma/sqrt(var+eps) per row (o*groups+g); gamma == nullptr: none
    const float* gamma;
    const float* var;
    float eps;
};

__device__ __forceinline__ float prescale_factor(const Prescale& ps, int64_t row) {
    // symbol/fold_bn_v1_gdrq.py:72  factor = bn_gamma / sqrt(bn_var + eps); each step rounded
    return __fdiv_rn(ps.gamma[row], __fsqrt_rn(__fadd_rn(ps.var[row], ps.eps)));
}

// VEC: 1 / 4 / 8 = register-staged loads of that many floats; 16 / 17 = the TMA-staged ring above (17: float -> double
// conversions on the integer pipe), `nst` stages of dynamic shared memory; 2 / 3 = large activations with SHORT rows that
// are not a multiple of eight floats (14x14 / 7x7 maps): the piece's (row, element) space flattened into one index,
// walked incrementally, four 128-bit (2) or eight scalar (3) loads in flight per thread
template <bool IS_MAX, int VEC>
__global__ void __launch_bounds__((VEC >= 8 || VEC == 2 || VEC == 3) ? B2Q_THREADS : 128)
reduce_seg_kernel(const float* __restrict__ x, SegPlan pl, Prescale ps, b2q_slot* slot, UpdateArgs u, int nst) {
    b2q_pdl_sync();
    __shared__ double smem[32];
    __shared__ unsigned int s_ticket;
    const SegPiece pc = seg_piece(pl);
    double acc = 0.0;
    float mx = 0.f;
    if (VEC >= 16) {
        extern __shared__ __align__(128) unsigned char s_dyn[];
        float* s_buf = reinterpret_cast<float*>(s_dyn);
        unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_dyn + (size_t)nst * B2Q_STREAM_CH * 4);
        double unused = 0.0;
        seg_stream_accumulate<IS_MAX ? 1 : 0, VEC == 17>(x, pl, pc, s_buf, s_bar, nst, acc, unused, mx);
    }
    if (VEC == 2 || VEC == 3) {
        constexpr int V = (VEC == 2) ? 4 : 1;
        constexpr int U = (VEC == 2) ? 4 : 8;
        const unsigned len = (unsigned)((pc.i1 - pc.i0) / V);
        const unsigned long long total = (unsigned long long)(pc.o1 - pc.o0) * len;
        const float* base = x + (pc.o0 * pl.groups + pc.g) * pl.inner + pc.i0;
        const int64_t ostride = pl.groups * pl.inner;
        const unsigned qT = blockDim.x / len, rT = blockDim.x % len;
        unsigned o = threadIdx.x / len, i = threadIdx.x % len;
        for (unsigned long long t = threadIdx.x; t < total; t += (unsigned long long)U * blockDim.x) {
            float v[U][V];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const float* p = base + (int64_t)o * ostride + (int64_t)i * V;
                if (t + (unsigned long long)k * blockDim.x < total) {
                    if (V == 4) {
                        const float4 t4 = *reinterpret_cast<const float4*>(p);
                        v[k][0] = t4.x; v[k][V > 1 ? 1 : 0] = t4.y; v[k][V > 2 ? 2 : 0] = t4.z; v[k][V > 3 ? 3 : 0] = t4.w;
                    } else {
                        v[k][0] = *p;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < V; ++e) v[k][e] = 0.f;
                }
                i += rT; o += qT;
                if (i >= len) { i -= len; ++o; }
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
#pragma unroll
                for (int e = 0; e < V; ++e) acc1<IS_MAX>(acc, mx, v[k][e]);
            }
        }
    }
    if (VEC == 8) {
        // 256-bit loads, four in flight per thread, over the piece's (row, word) space FLATTENED into one index: short
        // rows (gs = 1 on 56x56 maps: 392 words per row) would otherwise leave most lanes of every iteration idle
        const unsigned wpr = (unsigned)((pc.i1 - pc.i0) >> 3);                    // words per row in this piece
        const unsigned total = (unsigned)(pc.o1 - pc.o0) * wpr;
        for (unsigned w0 = threadIdx.x; w0 < total; w0 += 4 * blockDim.x) {
            f8 v[4];
            float fk[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned w = w0 + k * blockDim.x;
                if (w < total) {
                    const unsigned o = w / wpr, i = w - o * wpr;
                    const int64_t row = (pc.o0 + o) * pl.groups + pc.g;
                    v[k] = ld_f8<0>(x + row * pl.inner + pc.i0 + 8 * (int64_t)i);
                    fk[k] = ps.gamma ? prescale_factor(ps, row) : 1.f;
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[k].v[e] = 0.f;
                    fk[k] = 1.f;
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (ps.gamma) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[k].v[e] = __fmul_rn(v[k].v[e], fk[k]);
                }
                acc8<IS_MAX>(acc, mx, v[k]);
            }
        }
    }
    for (int64_t o = pc.o0; (VEC == 1 || VEC == 4) && o < pc.o1; ++o) {
        const int64_t row = o * pl.groups + pc.g;
        const float* base = x + row * pl.inner;
        const float f = ps.gamma ? prescale_factor(ps, row) : 1.f;
        if (VEC == 4) {
            const float4* b4 = reinterpret_cast<const float4*>(base);
            const int64_t end = pc.i1 >> 2;
            for (int64_t i = (pc.i0 >> 2) + threadIdx.x; i < end; i += 4 * (int64_t)blockDim.x) {
                float4 v[4];   // four independent 128-bit loads in flight per thread
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int64_t j = i + (int64_t)k * blockDim.x;
                    v[k] = (j < end) ? b4[j] : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (ps.gamma) { v[k].x = __fmul_rn(v[k].x, f); v[k].y = __fmul_rn(v[k].y, f); v[k].z = __fmul_rn(v[k].z, f); v[k].w = __fmul_rn(v[k].w, f); }
                    acc1<IS_MAX>(acc, mx, v[k].x); acc1<IS_MAX>(acc, mx, v[k].y);
                    acc1<IS_MAX>(acc, mx, v[k].z); acc1<IS_MAX>(acc, mx, v[k].w);
                }
            }
        } else {
            for (int64_t i = pc.i0 + threadIdx.x; i < pc.i1; i += blockDim.x) {
                float v = base[i];
                if (ps.gamma) v = __fmul_rn(v, f);
                acc1<IS_MAX>(acc, mx, v);
            }
        }
    }
    double r = block_reduce<IS_MAX>(IS_MAX ? (double)mx : acc, smem);
    const int SP = pl.S * pl.P;
    const float count = (float)(pl.outer * pl.inner);  // elements per group (float32(N) like MXNet's mean)
    if (SP == 1) {   // one piece per group: this block owns the whole group, nothing to combine
        if (threadIdx.x == 0) apply_update(u, (int)pc.g, IS_MAX ? (float)r : __fdiv_rn((float)r, count));
        return;
    }
    if (threadIdx.x == 0) {
        slot->partial[blockIdx.x] = r;
        s_ticket = b2q_take_ticket(&slot->ticket, gridDim.x - 1);
    }
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    if (pl.groups >= 64) {
        // many groups: one thread per group walks its SP partials in order (fixed order => deterministic)
        for (int64_t gg = threadIdx.x; gg < pl.groups; gg += blockDim.x) {
            double a = 0.0;
            float m = 0.f;
            for (int l = 0; l < SP; ++l) {
                const double q = __ldcg(&slot->partial[gg * SP + l]);
                if (IS_MAX) m = fmax_nan(m, (float)q); else a += q;
            }
            apply_update(u, (int)gg, IS_MAX ? m : __fdiv_rn((float)a, count));
        }
    } else {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
        for (int64_t gg = wid; gg < pl.groups; gg += nw) {
            double a = 0.0;
            float m = 0.f;
            for (int l = lane; l < SP; l += 32) {
                double q = __ldcg(&slot->partial[gg * SP + l]);
                if (IS_MAX) m = fmax_nan(m, (float)q); else a += q;
            }
            if (IS_MAX) m = warp_max(m); else a = warp_sum(a);
            if (lane == 0) apply_update(u, (int)gg, IS_MAX ? m : __fdiv_rn((float)a, count));
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) slot->ticket = 0;
}

// Stand-alone K3 (after a cross-rank allreduce of the statistic).
static __global__ void threshold_update_kernel(const float* __restrict__ stat, int groups, UpdateArgs u) {
    b2q_pdl_sync();
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < groups) apply_update(u, g, stat[g]);
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
static inline int64_t b2q_flat_grid(const b2q_ctx* ctx, int64_t n8, int unroll, int bps = 0) {
    const int64_t tile = (int64_t)B2Q_THREADS * unroll;
    int64_t ntiles = (n8 + tile - 1) / tile;
    int64_t grid = (int64_t)ctx->num_sms * (bps > 0 ? bps : ctx->blocks_per_sm);
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) grid = 1;
    if (bps > 0 && grid > B2Q_MAX_PIECES) grid = B2Q_MAX_PIECES;   // reductions keep one partial per block
    if (grid > 0x7fffffff) grid = 0x7fffffff;
    return grid;
}

// TMA-staged reductions: eligibility (rows of a piece are contiguous runs of >= 1 KB, 16-byte aligned -- guaranteed by the
// callers' 256-bit checks), number of stages and dynamic shared memory
static inline bool b2q_stream_ok(const SegPlan& pl) {
    const int64_t len = pl.P > 1 ? pl.part : pl.inner;
    const int64_t last = pl.inner - (int64_t)(pl.P - 1) * pl.part;     // the last inner part may be shorter
    return len >= 256 && last >= 256 && (len % 8) == 0 && (last % 8) == 0;
}

static inline int b2q_stream_stages(const b2q_ctx* ctx) {
    int n = ctx->stream_stages;
    if (n < 2) n = 2;
    if (n > 12) n = 12;
    return n;
}

static inline size_t b2q_stream_smem(int nst) { return (size_t)nst * B2Q_STREAM_CH * 4 + (size_t)nst * 8; }

template <typename K>
static inline cudaError_t b2q_allow_smem(K kernel, size_t smem) { return b2q_kernel_smem_once((const void*)kernel, smem); }

// Reduction half of the fused whole-tensor forward with the update deferred to the consumer.  Returns the number of
// partials through *n_partials (0: tensor not eligible, caller must use launch_reduce).
template <bool IS_MAX>
static int launch_reduce_deferred(b2q_ctx* ctx, b2q_slot* slot, const float* x, int64_t n, UpdateArgs u,
                                  cudaStream_t st, int* n_partials) {
    *n_partials = 0;
    FlatSplit sp = b2q_flat_split(x, n);
    if (sp.head > B2Q_THREADS) return 0;
    const int64_t grid = b2q_flat_grid(ctx, sp.n8, B2Q_REDUCE_UNROLL,
                                       IS_MAX ? ctx->reduce_deferred_blocks_per_sm : ctx->reduce_blocks_per_sm);
    b2q_timed_launch tl(ctx, B2Q_KIND_REDUCE_FLAT, 4.0 * (double)n, st);
    b2q_launch(ctx, reduce_flat_kernel<IS_MAX, B2Q_REDUCE_UNROLL, B2Q_REDUCE_LDPOL, false>, (unsigned)grid, B2Q_THREADS, st,
               x, sp, slot, u, (float)n);
    B2Q_LAUNCH_CHECK(ctx);
    *n_partials = (int)grid;
    return 0;
}

template <bool IS_MAX>
static int launch_reduce(b2q_ctx* ctx, b2q_slot* slot, const float* x, int64_t outer, int64_t groups,
                         int64_t inner, Prescale ps, UpdateArgs u, cudaStream_t st) {
    B2Q_REQUIRE(outer >= 1 && groups >= 1 && inner >= 1, "empty tensor in reduction");
    B2Q_REQUIRE(groups <= B2Q_MAX_GROUPS, "too many groups/channels (max 8192)");
    const int64_t n = outer * groups * inner;
    if (groups == 1 && ps.gamma == nullptr) {
        FlatSplit sp = b2q_flat_split(x, n);
        if (sp.head <= B2Q_THREADS) {
            const int64_t grid = b2q_flat_grid(ctx, sp.n8, B2Q_REDUCE_UNROLL, ctx->reduce_blocks_per_sm);
            b2q_timed_launch tl(ctx, B2Q_KIND_REDUCE_FLAT, 4.0 * (double)n, st);
            b2q_launch(ctx, reduce_flat_kernel<IS_MAX, B2Q_REDUCE_UNROLL, B2Q_REDUCE_LDPOL, true>, (unsigned)grid,
                       B2Q_THREADS, st, x, sp, slot, u, (float)n);
            B2Q_LAUNCH_CHECK(ctx);
            return 0;
        }
        outer = 1; inner = n;  // not even float-aligned: scalar segmented path
    }
    SegPlan pl = b2q_seg_plan(x, nullptr, outer, groups, inner, ctx->num_sms * 16);
    const unsigned grid = (unsigned)(groups * pl.S * pl.P);
    b2q_timed_launch tl(ctx, B2Q_KIND_OTHER, 4.0 * (double)n, st);
    const bool vec8 = pl.vec == 4 && inner % 8 == 0 && pl.part % 8 == 0 && (((uintptr_t)x) & 31) == 0 && n >= (1 << 20);
    if (vec8 && ctx->stream_reduce && ps.gamma == nullptr && b2q_stream_ok(pl)) {
        // large activations with rows of at least 1 KB per piece: TMA-staged ring (see seg_stream_accumulate)
        const int nst = b2q_stream_stages(ctx);
        const size_t smem = b2q_stream_smem(nst);
        if (ctx->stream_icvt) {
            B2Q_CHECK_CUDA(b2q_allow_smem(reduce_seg_kernel<IS_MAX, 17>, smem));
            b2q_launch_smem(ctx, reduce_seg_kernel<IS_MAX, 17>, grid, B2Q_THREADS, smem, st, x, pl, ps, slot, u, nst);
        } else {
            B2Q_CHECK_CUDA(b2q_allow_smem(reduce_seg_kernel<IS_MAX, 16>, smem));
            b2q_launch_smem(ctx, reduce_seg_kernel<IS_MAX, 16>, grid, B2Q_THREADS, smem, st, x, pl, ps, slot, u, nst);
        }
    } else if (vec8)
        b2q_launch(ctx, reduce_seg_kernel<IS_MAX, 8>, grid, B2Q_THREADS, st, x, pl, ps, slot, u, 0);
    else if (ps.gamma == nullptr && pl.P == 1 && inner < 2048 && n >= (1 << 18)) {   // short rows: flattened walk
        if (pl.vec == 4) b2q_launch(ctx, reduce_seg_kernel<IS_MAX, 2>, grid, B2Q_THREADS, st, x, pl, ps, slot, u, 0);
        else b2q_launch(ctx, reduce_seg_kernel<IS_MAX, 3>, grid, B2Q_THREADS, st, x, pl, ps, slot, u, 0);
    }
    else if (pl.vec == 4) b2q_launch(ctx, reduce_seg_kernel<IS_MAX, 4>, grid, 128, st, x, pl, ps, slot, u, 0);
    else b2q_launch(ctx, reduce_seg_kernel<IS_MAX, 1>, grid, 128, st, x, pl, ps, slot, u, 0);
    B2Q_LAUNCH_CHECK(ctx);
    return 0;
}
