// the step either side of the fold-BN operator.  In the reference's graph
// (symbol/fold_bn_v1_gdrq.py:268-287) a BatchNorm_v1(output_mean_var=True) reduces the convolution output to its batch
// mean / variance, and GDRQ_Fold_BN then folds gamma / sqrt(var + eps) into the weight, quantises it per out-channel and
// builds the folded bias (:70-96,113): a per-channel reduction over the (N, C, H, W) activation followed by a chain of
// small per-channel launches.  Here it is ONE kernel: every block reduces one piece of one channel (sum and sum of
// squares in fp64); the block that completes a channel (per-channel release ticket) turns the partials into mean / var
// and immediately folds, clips and quantises that channel's weight row and writes its bias -- no second launch, no
// per-channel host round trip, the statistics never leave the chip between the two steps.
//
// Numerics of the statistics [upstream src/operator/batch_norm_v1-inl.h]: scale = fl(C / size);
// mean = fl(scale * sum(x)); var = fl(scale * sum((x - mean)^2)), sums rounded once.  The squared deviations are
// accumulated here as SS - 2 mean S + n mean^2 in fp64 (mean = the ROUNDED float32 mean, as the reference subtracts it),
// which agrees with the reference's fp32 (x - mean)^2 terms to ~1e-7 relative: mean / var carry the 1e-6 tolerance of
// every mean-derived quantity (BASELINE.json north_star); given equal mean / var the folded, quantised weight and the
// bias are bit-exact (tests/test_gpu_bnfold.py).
#include <cstring>

#include "b2q_common.cuh"
#include "b2q_qdq.cuh"
#include "b2q_reduce.cuh"

#define B2Q_CTX(ctx)                               \
    B2Q_REQUIRE((ctx) != nullptr, "null context"); \
    B2Q_CHECK_CUDA(cudaSetDevice((ctx)->device))

struct BnFold {
    const float* w;        // [C, cols] or null: statistics only
    float* w_q;
    float* bias;
    float* aux_w;          // [C]
    const float* gamma;
    const float* beta;
    float eps;
    int64_t cols;
    int is_train, fast;
};

// Register budget matters more than anything else here: the first version held four 256-bit words plus their
// double-precision images per thread (94 registers -> 2 blocks per SM, 24 % of the warps, 30 % of DRAM throughput in ncu,
// profiles/r02g_ncu_micro.csv; 0.38 of the copy peak).  Now each element is converted, added and squared inside ONE asm
// statement (acc1_sq_asm: the compiler cannot convert a batch of words first and keep 32 doubles alive), two alternating
// chains per sum: 48 registers with two loads in flight and five resident blocks -> 0.72-0.78 of the copy peak
// (profiles/r02i_bnstat_pipes.md; the segmented max|x| reduction, which has no fp64 work at all, reaches 0.82 on the
// same access pattern).  BN_UNROLL / BN_BLOCKS / ICVT variants are selectable (options bn_variant, stream_icvt).
template <int VEC, int BN_UNROLL, int BN_BLOCKS, int ICVT>
__global__ void __launch_bounds__(B2Q_THREADS, BN_BLOCKS)
bnstat_fold_kernel(const float* __restrict__ y, SegPlan pl, b2q_slot* slot, float scale, float* __restrict__ mean_out,
                   float* __restrict__ var_out, BnFold f, int nst) {
    b2q_pdl_sync();
    __shared__ double smem[32];
    __shared__ unsigned int s_ticket;
    __shared__ float s_val[2];
    const SegPiece pc = seg_piece(pl);
    double s = 0.0, ss = 0.0;
    if (VEC >= 16) {   // TMA-staged ring (b2q_reduce.cuh: seg_stream_accumulate); 17: conversions on the integer pipe
        extern __shared__ __align__(128) unsigned char s_dyn[];
        float* s_buf = reinterpret_cast<float*>(s_dyn);
        unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_dyn + (size_t)nst * B2Q_STREAM_CH * 4);
        float unused = 0.f;
        seg_stream_accumulate<2, VEC == 17>(y, pl, pc, s_buf, s_bar, nst, s, ss, unused);
    } else if (VEC == 5) {   // rows that are not a multiple of eight floats: masked aligned 256-bit words (b2q_reduce.cuh)
        double s1 = 0.0, q0 = 0.0, q1 = 0.0;
        seg_masked_words<2>(y, pl.outer * pl.groups * pl.inner, pl, pc,
                            [&](const f8& w) { acc8_sq_seq<0>(s, s1, q0, q1, w); });
        s += s1;
        ss = q0 + q1;
    } else if (VEC == 8) {
        const unsigned wpr = (unsigned)((pc.i1 - pc.i0) >> 3);
        const unsigned total = (unsigned)(pc.o1 - pc.o0) * wpr;
        const float* base = y + (pc.o0 * pl.groups + pc.g) * pl.inner + pc.i0;
        const int64_t ostride = pl.groups * pl.inner;
        double q0 = 0.0, q1 = 0.0, s1 = 0.0;
        for (unsigned w0 = threadIdx.x; w0 < total; w0 += BN_UNROLL * blockDim.x) {
            f8 v[BN_UNROLL];
#pragma unroll
            for (int k = 0; k < BN_UNROLL; ++k) {
                const unsigned w = w0 + k * blockDim.x;
                if (w < total) {
                    const unsigned o = w / wpr, i = w - o * wpr;
                    v[k] = ld_f8<1>(base + o * ostride + 8 * (int64_t)i);
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[k].v[e] = 0.f;
                }
            }
#pragma unroll
            for (int k = 0; k < BN_UNROLL; ++k) {
                acc8_sq_seq<ICVT>(s, s1, q0, q1, v[k]);
            }
        }
        s += s1;
        ss = q0 + q1;
    } else {
        // Feature maps whose rows are not a multiple of eight floats -- 14x14 (196) and 7x7 (49), half the layers of
        // ResNet / MobileNet: the piece's (row, element) space flattened into ONE index, so that a row of 49 floats does
        // not leave 207 of 256 threads idle, walked incrementally (no division per element), four 128-bit loads (rows a
        // multiple of four floats, 16-byte aligned) or sixteen scalar loads in flight per thread.  See
        // profiles/r02v_bnstat_small_maps.md (0.14 -> 0.57 on 14x14, 0.036 -> 0.26 on 7x7)
        constexpr int V = (VEC == 4) ? 4 : 1;
        constexpr int U = (VEC == 4) ? 4 : 16;
        const unsigned len = (unsigned)((pc.i1 - pc.i0) / V);          // vectors per row inside this piece
        const unsigned long long total = (unsigned long long)(pc.o1 - pc.o0) * len;
        const float* base = y + (pc.o0 * pl.groups + pc.g) * pl.inner + pc.i0;
        const int64_t ostride = pl.groups * pl.inner;
        const unsigned T = blockDim.x;
        const unsigned qT = T / len, rT = T % len;
        unsigned o = threadIdx.x / len, i = threadIdx.x % len;
        double s1 = 0.0, q0 = 0.0, q1 = 0.0;
        for (unsigned long long w = threadIdx.x; w < total; w += (unsigned long long)U * T) {
            float v[U][V];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const float* p = base + (int64_t)o * ostride + (int64_t)i * V;
                if (w + (unsigned long long)k * T < total) {
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
                for (int e = 0; e < V; ++e) {
                    if ((k * V + e) & 1) acc1_sq_asm(s1, q1, v[k][e]); else acc1_sq_asm(s, q0, v[k][e]);
                }
            }
        }
        s += s1;
        ss = q0 + q1;
    }
    const double bs = block_reduce<false>(s, smem);
    const double bss = block_reduce<false>(ss, smem);
    const int SP = pl.S * pl.P;
    const int c = (int)pc.g;
    double S = bs, SS = bss;     // one piece per channel (small feature maps with many channels): this block owns the
    if (SP > 1) {                // channel -- no partials, no ticket, no second reduction
        if (threadIdx.x == 0) {
            slot->partial[2 * blockIdx.x] = bs;
            slot->partial[2 * blockIdx.x + 1] = bss;
            s_ticket = b2q_take_ticket(&slot->row_ticket[c], (unsigned)SP - 1u);
        }
        __syncthreads();
        if (s_ticket != (unsigned)SP - 1u) return;
        // ---- this block completes channel c: statistics in a fixed order ----
        double a = 0.0, b = 0.0;
        for (int l = threadIdx.x; l < SP; l += blockDim.x) {
            a += __ldcg(&slot->partial[2 * ((int64_t)c * SP + l)]);
            b += __ldcg(&slot->partial[2 * ((int64_t)c * SP + l) + 1]);
        }
        S = block_reduce<false>(a, smem);
        SS = block_reduce<false>(b, smem);
    }
    if (threadIdx.x == 0) {
        const double n = (double)(pl.outer * pl.inner);
        const float mean = __fmul_rn(scale, (float)S);
        const double dm = (double)mean;
        const double dev = SS - 2.0 * dm * S + n * dm * dm;          // sum of (x - mean)^2
        s_val[0] = mean;
        s_val[1] = __fmul_rn(scale, (float)(dev < 0.0 ? 0.0 : dev));   // NaN (a NaN / Inf in the channel) stays NaN
        mean_out[c] = s_val[0];
        var_out[c] = s_val[1];
        if (SP > 1) slot->row_ticket[c] = 0;
    }
    __syncthreads();
    if (f.w == nullptr) return;
    // ---- fold-BN weight path of channel c (symbol/fold_bn_v1_gdrq.py:70-96,113), same arithmetic as rows_*_kernel ----
    const float mean = s_val[0], var = s_val[1];
    const float den = __fsqrt_rn(__fadd_rn(var, f.eps));
    const float factor = __fdiv_rn(f.gamma[c], den);                  // :72
    const float* wr = f.w + (int64_t)c * f.cols;
    float* qr = f.w_q + (int64_t)c * f.cols;
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < f.cols; i += blockDim.x) acc += (double)fabsf(__fmul_rn(wr[i], factor));
    const double tot = block_reduce<false>(acc, smem);
    if (threadIdx.x == 0) {
        const float T = __fmul_rn(2.f, __fdiv_rn((float)tot, (float)f.cols));   // :82
        s_val[0] = T;
        if (f.is_train && f.aux_w) f.aux_w[c] = T;                               // :94-95
        if (f.bias) f.bias[c] = __fsub_rn(f.beta[c], __fdiv_rn(__fmul_rn(mean, f.gamma[c]), den));   // :113
    }
    __syncthreads();
    const float T = s_val[0];
    const QScale qs = make_qscale(T, 127.f, f.fast != 0);
    for (int64_t i = threadIdx.x; i < f.cols; i += blockDim.x) {
        const float v = mx_clip(__fmul_rn(wr[i], factor), -T, T);
        qr[i] = __fmul_rn(quant_code(v, qs), qs.q);
    }
}

static int launch_bnstat(b2q_ctx* ctx, const float* y, int64_t n, int64_t c, int64_t hw, float* mean, float* var, BnFold f,
                         cudaStream_t st) {
    B2Q_REQUIRE(y && mean && var && n >= 1 && c >= 1 && hw >= 1, "bad argument");
    B2Q_REQUIRE(c <= B2Q_MAX_GROUPS, "too many channels (max 8192)");
    SegPlan pl = b2q_seg_plan(y, nullptr, n, c, hw, ctx->num_sms * (ctx->bn_pieces_per_sm > 0 ? ctx->bn_pieces_per_sm : 16));
    // many small channels (7x7 / 14x14 maps with >= 4 channels per SM, <= 256 KB each): one block per channel, which then
    // needs no partials, ticket or second reduction
    if (c >= 4 * (int64_t)ctx->num_sms && n * hw * 4 <= (256 << 10) && pl.P == 1) pl.S = 1;
    while ((int64_t)c * pl.S * pl.P > B2Q_MAX_PIECES / 2 && pl.S > 1) --pl.S;   // two partials per piece
    B2Q_REQUIRE((int64_t)c * pl.S * pl.P <= B2Q_MAX_PIECES / 2, "activation too large for one statistics launch");
    // fl(C / size) as batch_norm_v1-inl.h computes it: two float operands
    const float scale = (float)c / (float)((double)n * (double)c * (double)hw);
    const unsigned grid = (unsigned)(c * pl.S * pl.P);
    b2q_slot* slot = b2q_take_slot(ctx, st);
    b2q_timed_launch tl(ctx, B2Q_KIND_OTHER, 4.0 * (double)(n * c * hw), st);
    const bool vec8 = (hw % 8 == 0) && (pl.part % 8 == 0) && ((((uintptr_t)y) & 31) == 0);
    const bool vec4 = !vec8 && (hw % 4 == 0) && (pl.part % 4 == 0) && ((((uintptr_t)y) & 15) == 0);
    const bool masked = !vec8 && !vec4 && ctx->seg_masked && pl.P == 1 && hw >= 8 && hw < 2048 && ((((uintptr_t)y) & 31) == 0) &&
                        n * c * hw >= (1 << 18);
    if (vec8 && ctx->stream_reduce && b2q_stream_ok(pl)) {
        const int nst = b2q_stream_stages(ctx);
        const size_t smem = b2q_stream_smem(nst);
        if (ctx->stream_icvt) {
            B2Q_CHECK_CUDA(b2q_allow_smem(bnstat_fold_kernel<17, 1, 1, 0>, smem));
            b2q_launch_smem(ctx, bnstat_fold_kernel<17, 1, 1, 0>, grid, B2Q_THREADS, smem, st, y, pl, slot, scale, mean, var, f, nst);
        } else {
            B2Q_CHECK_CUDA(b2q_allow_smem(bnstat_fold_kernel<16, 1, 1, 0>, smem));
            b2q_launch_smem(ctx, bnstat_fold_kernel<16, 1, 1, 0>, grid, B2Q_THREADS, smem, st, y, pl, slot, scale, mean, var, f, nst);
        }
        B2Q_LAUNCH_CHECK(ctx);
        return 0;
    }
    // (loads in flight per thread, resident blocks per SM); option bn_variant
#define B2Q_BN_LAUNCH(U, B) do {                                                                                          \
        if (masked) b2q_launch(ctx, bnstat_fold_kernel<5, U, B, 0>, grid, B2Q_THREADS, st, y, pl, slot, scale, mean, var, f, 0);  \
        else if (vec4) b2q_launch(ctx, bnstat_fold_kernel<4, U, B, 0>, grid, B2Q_THREADS, st, y, pl, slot, scale, mean, var, f, 0); \
        else if (!vec8) b2q_launch(ctx, bnstat_fold_kernel<1, U, B, 0>, grid, B2Q_THREADS, st, y, pl, slot, scale, mean, var, f, 0);  \
        else if (ctx->stream_icvt == 1)                                                                                    \
            b2q_launch(ctx, bnstat_fold_kernel<8, U, B, 1>, grid, B2Q_THREADS, st, y, pl, slot, scale, mean, var, f, 0);       \
        else if (ctx->stream_icvt == 2)                                                                                    \
            b2q_launch(ctx, bnstat_fold_kernel<8, U, B, 2>, grid, B2Q_THREADS, st, y, pl, slot, scale, mean, var, f, 0);       \
        else b2q_launch(ctx, bnstat_fold_kernel<8, U, B, 0>, grid, B2Q_THREADS, st, y, pl, slot, scale, mean, var, f, 0);      \
    } while (0)
    switch (ctx->bn_variant) {
        case 1: B2Q_BN_LAUNCH(2, 5); break;
        case 2: B2Q_BN_LAUNCH(4, 4); break;
        case 3: B2Q_BN_LAUNCH(4, 5); break;
        case 4: B2Q_BN_LAUNCH(2, 8); break;
        case 5: B2Q_BN_LAUNCH(4, 6); break;
        case 6: B2Q_BN_LAUNCH(2, 6); break;
        default: B2Q_BN_LAUNCH(2, 4); break;
    }
#undef B2Q_BN_LAUNCH
    B2Q_LAUNCH_CHECK(ctx);
    return 0;
}

extern "C" {

int b2q_bn_batch_stats_f32(b2q_ctx* ctx, const float* y, int64_t n, int64_t c, int64_t hw, float* mean, float* var,
                           void* stream) {
    B2Q_CTX(ctx);
    BnFold none;
    memset(&none, 0, sizeof(none));
    return launch_bnstat(ctx, y, n, c, hw, mean, var, none, (cudaStream_t)stream);
}

int b2q_bnstat_foldbn_weight_fwd_f32(b2q_ctx* ctx, const float* conv_out, int64_t n, int64_t c, int64_t hw, float* mean,
                                     float* var, const float* w, float* w_q, float* bias, float* aux_weight,
                                     const float* gamma, const float* beta, float eps, int64_t cols, int per_channel,
                                     int quantize, int is_train, void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(w && w_q && bias && gamma && beta && cols >= 1, "bad argument");
    if (per_channel && quantize) {
        B2Q_REQUIRE(aux_weight != nullptr, "null aux");
        BnFold f = {w, w_q, bias, aux_weight, gamma, beta, eps, cols, is_train, ctx->fast_div};
        return launch_bnstat(ctx, conv_out, n, c, hw, mean, var, f, (cudaStream_t)stream);
    }
    // per-tensor threshold (needs every channel's statistics) or no quantisation: statistics, then the weight kernels
    BnFold none;
    memset(&none, 0, sizeof(none));
    int rc = launch_bnstat(ctx, conv_out, n, c, hw, mean, var, none, (cudaStream_t)stream);
    if (rc) return rc;
    return b2q_foldbn_weight_fwd_f32(ctx, w, w_q, bias, aux_weight, gamma, beta, mean, var, eps, c, cols, per_channel,
                                     quantize, is_train, stream);
}

}  // extern "C"
