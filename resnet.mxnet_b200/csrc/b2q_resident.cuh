// Single-launch forward for tensors that fit on chip (SURVEY.md K1+K3+K4 in one kernel).
//
// A 26 MB or 51 MB activation is read from HBM by the reduction and, a kernel later, read again by the QDQ sweep.  Here
// one persistent grid (one 512-thread CTA per SM, ~216 KB of shared memory each = 32 MB across the GPU) does both:
//
//   phase 1   every CTA owns a contiguous chunk.  One elected thread copies the first <= 216 KB of it into shared memory
//             with bulk asynchronous copies (TMA engine, one mbarrier counting the bytes); meanwhile the threads reduce
//             the rest of the chunk straight from global memory (plain loads: the lines stay in L2), then the staged
//             part out of shared memory.  One tagged atomicMax (max) or one fp64 partial (mean) per CTA.
//   barrier   grid-wide: release ticket; the last CTA to arrive publishes the tag, the others acquire it.  All CTAs are
//             resident by construction (grid <= SM count, shared memory allows one CTA per SM).
//   phase 2   every CTA derives the threshold (EMA / first batch / GDRQ alpha / 2*mean) in registers exactly like the
//             two-kernel path, CTA 0 writes aux, and the chunk is quantised: the staged part from shared memory, the
//             rest from L2.  x crosses HBM once instead of twice; launch + dependency latency is paid once.
//
// Results are bit-identical to the two-kernel path (same reduction semantics, same qdq8).  Algorithmic bytes stay
// 12 B/element for the roofline ("8 if the single-read variant applies -- still report against 12").
#pragma once
#include "b2q_common.cuh"
#include "b2q_qdq.cuh"
#include "b2q_reduce.cuh"

#define B2Q_RES_THREADS 512
#define B2Q_RES_SMEM_BYTES (216 * 1024)
#define B2Q_RES_PIECE_BYTES (16 * 1024)

__device__ __forceinline__ void st_release_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <bool IS_MAX, int CLIP>
__global__ void __launch_bounds__(B2Q_RES_THREADS, 1)
fused_resident_kernel(const float* __restrict__ x, float* __restrict__ y, FlatSplit sp, b2q_slot* slot, UpdateArgs u,
                      float qlevel, int fast, int clip_with_fresh, float count, int smem_words) {
    extern __shared__ __align__(128) float s_stage[];
    __shared__ double s_red[32];
    __shared__ float s_stat;
    __shared__ unsigned int s_tag;
    __shared__ __align__(8) unsigned long long s_bar;
    b2q_pdl_sync();
    const int tid = threadIdx.x;
    const int64_t W = (sp.n8 + gridDim.x - 1) / gridDim.x;          // 256-bit words per CTA
    const int64_t w0 = (int64_t)blockIdx.x * W;
    int64_t nw = sp.n8 - w0;
    if (nw > W) nw = W;
    if (nw < 0) nw = 0;
    const int64_t ns = nw < (int64_t)smem_words ? nw : (int64_t)smem_words;   // staged words
    const float* xb = x + sp.head + 8 * w0;
    float* yb = y + sp.head + 8 * w0;
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
        s_tag = slot->epoch + 1u;       // strictly newer than anything the slot's words hold
    }
    __syncthreads();
    if (tid == 0 && ns > 0) {
        const unsigned total = (unsigned)ns * 32u;
        mbar_expect_tx(&s_bar, total);
        for (unsigned off = 0; off < total; off += B2Q_RES_PIECE_BYTES) {
            const unsigned bytes = total - off < B2Q_RES_PIECE_BYTES ? total - off : B2Q_RES_PIECE_BYTES;
            bulk_g2s(reinterpret_cast<char*>(s_stage) + off, reinterpret_cast<const char*>(xb) + off, bytes, &s_bar);
        }
    }
    // ---- phase 1: statistic ----
    double acc = 0.0;
    float mx = 0.f;
    for (int64_t i0 = ns + tid; i0 < nw; i0 += 4 * B2Q_RES_THREADS) {       // the part that does not fit: from global
        f8 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t i = i0 + (int64_t)k * B2Q_RES_THREADS;
            if (i < nw) v[k] = ld_f8<0>(xb + 8 * i);
            else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[k].v[j] = 0.f;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) acc8<IS_MAX>(acc, mx, v[k]);
    }
    if (blockIdx.x == 0) {  // the (at most 14) unaligned scalars
        if ((int64_t)tid < sp.head) acc1<IS_MAX>(acc, mx, x[tid]);
        if ((int64_t)tid < sp.tail) acc1<IS_MAX>(acc, mx, x[sp.head + 8 * sp.n8 + tid]);
    }
    const float4* s4 = reinterpret_cast<const float4*>(s_stage);
    const int n4 = (int)ns * 2;
    if (ns > 0) mbar_wait(&s_bar, 0);
    for (int i0 = tid; i0 < n4; i0 += 4 * B2Q_RES_THREADS) {               // consecutive threads, consecutive 16 bytes
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = i0 + k * B2Q_RES_THREADS;
            v[k] = i < n4 ? s4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        f8 a, b;
        a.v[0] = v[0].x; a.v[1] = v[0].y; a.v[2] = v[0].z; a.v[3] = v[0].w;
        a.v[4] = v[1].x; a.v[5] = v[1].y; a.v[6] = v[1].z; a.v[7] = v[1].w;
        b.v[0] = v[2].x; b.v[1] = v[2].y; b.v[2] = v[2].z; b.v[3] = v[2].w;
        b.v[4] = v[3].x; b.v[5] = v[3].y; b.v[6] = v[3].z; b.v[7] = v[3].w;
        acc8<IS_MAX>(acc, mx, a);
        acc8<IS_MAX>(acc, mx, b);
    }
    const double r = block_reduce<IS_MAX>(IS_MAX ? (double)mx : acc, s_red);
    // ---- grid-wide barrier ----
    if (tid == 0) {
        const unsigned int tag = s_tag;
        if (IS_MAX) atomicMax(&slot->max64, ((unsigned long long)tag << 32) | __float_as_uint((float)r));
        else slot->partial[blockIdx.x] = r;
        if (blockIdx.x == 0 && u.aux) slot->scale[0] = u.aux[0];     // snapshot of the old threshold
        const unsigned int t = b2q_take_ticket(&slot->ticket, gridDim.x - 1);
        if (t == gridDim.x - 1) {
            slot->ticket = 0;
            st_release_u32(&slot->epoch, tag);                        // everybody's contribution is visible: go
        } else {
            while (ld_acquire_u32(&slot->epoch) != tag) __nanosleep(32);
        }
    }
    __syncthreads();
    // ---- threshold, in registers ----
    float stat;
    if (IS_MAX) {
        stat = __uint_as_float((unsigned int)(__ldcg(&slot->max64) & 0xffffffffull));
    } else {
        double p = (tid < (int)gridDim.x) ? __ldcg(&slot->partial[tid]) : 0.0;   // grid <= 512: one partial per thread
        const double tot = block_reduce<false>(p, s_red);                        // the same fixed tree in every CTA
        if (tid == 0) s_stat = __fdiv_rn((float)tot, count);
        __syncthreads();
        stat = s_stat;
    }
    const float a_old = u.aux ? __ldcg(&slot->scale[0]) : 0.f;
    float fresh, next;
    compute_update(u.mode, u.p0, u.p1, a_old, stat, fresh, next);
    const float after = u.write_aux ? next : a_old;
    const float T = u.use_aux_as_scale ? after : fresh;
    const float Tc = clip_with_fresh ? fresh : T;
    if (blockIdx.x == 0 && tid == 0 && u.write_aux && u.aux) u.aux[0] = next;
    const bool needs_pos = (CLIP != B2Q_CLIP_NONE && CLIP != B2Q_CLIP_PACT);
    const QScale s = make_qscale(T, qlevel, fast != 0 && !(needs_pos && !(Tc >= 0.f)));
    // ---- phase 2: quantise-dequantise; staged part from shared memory, the rest from L2 ----
    float4* y4 = reinterpret_cast<float4*>(yb);
    for (int i0 = tid; i0 < n4; i0 += 2 * B2Q_RES_THREADS) {
        const int i1 = i0 + B2Q_RES_THREADS;
        const float4 va = s4[i0];
        const float4 vb = i1 < n4 ? s4[i1] : make_float4(0.f, 0.f, 0.f, 0.f);
        f8 in, o;
        in.v[0] = va.x; in.v[1] = va.y; in.v[2] = va.z; in.v[3] = va.w;
        in.v[4] = vb.x; in.v[5] = vb.y; in.v[6] = vb.z; in.v[7] = vb.w;
        qdq8<CLIP>(in, o, Tc, s);
        y4[i0] = make_float4(o.v[0], o.v[1], o.v[2], o.v[3]);
        if (i1 < n4) y4[i1] = make_float4(o.v[4], o.v[5], o.v[6], o.v[7]);
    }
    for (int64_t i0 = ns + tid; i0 < nw; i0 += 2 * B2Q_RES_THREADS) {
        const int64_t i1 = i0 + B2Q_RES_THREADS;
        f8 v0, v1, o;
        v0 = ld_f8<B2Q_QDQ_LDPOL>(xb + 8 * i0);
        if (i1 < nw) v1 = ld_f8<B2Q_QDQ_LDPOL>(xb + 8 * i1);
        qdq8<CLIP>(v0, o, Tc, s);
        st_f8<0>(yb + 8 * i0, o);
        if (i1 < nw) {
            qdq8<CLIP>(v1, o, Tc, s);
            st_f8<0>(yb + 8 * i1, o);
        }
    }
    if (blockIdx.x == 0) {  // unaligned head / tail scalars
        int64_t idx = -1;
        if ((int64_t)tid < sp.head) idx = tid;
        else if ((int64_t)tid - sp.head < sp.tail) idx = sp.head + 8 * sp.n8 + ((int64_t)tid - sp.head);
        if (idx >= 0) y[idx] = __fmul_rn(quant_code_exact(clip_value(CLIP, x[idx], Tc), s.q), s.q);
    }
}

// *done = 0: not eligible (caller takes the two-kernel path).
template <bool IS_MAX>
[[maybe_unused]] static int launch_fused_resident(b2q_ctx* ctx, b2q_slot* slot, const float* x, float* y, int64_t n,
                                                  UpdateArgs u, float qlevel, int clip_mode, int clip_with_fresh,
                                                  cudaStream_t st, int* done) {
    *done = 0;
    if (!ctx->resident || n * 4 > ((int64_t)ctx->resident_max_mb << 20) || n < (1 << 20)) return 0;
    if (!(clip_mode == B2Q_CLIP_NONE || clip_mode == B2Q_CLIP_SYM) || u.stat_out != nullptr) return 0;
    FlatSplit sp = b2q_flat_split(x, n);
    if (!same_misalignment(x, y) || sp.head > B2Q_THREADS || sp.n8 < 1) return 0;
    bool& optin = ctx->resident_optin[(IS_MAX ? 0 : 2) + (clip_mode == B2Q_CLIP_SYM ? 1 : 0)];
#define B2Q_RES_KERNEL(C) fused_resident_kernel<IS_MAX, C>
    if (!optin) {
        if (clip_mode == B2Q_CLIP_SYM)
            B2Q_CHECK_CUDA(cudaFuncSetAttribute(B2Q_RES_KERNEL(B2Q_CLIP_SYM), cudaFuncAttributeMaxDynamicSharedMemorySize, B2Q_RES_SMEM_BYTES));
        else
            B2Q_CHECK_CUDA(cudaFuncSetAttribute(B2Q_RES_KERNEL(B2Q_CLIP_NONE), cudaFuncAttributeMaxDynamicSharedMemorySize, B2Q_RES_SMEM_BYTES));
        optin = true;
    }
    // one CTA per SM; small tensors use fewer CTAs (at least 512 words = 16 KB each)
    int64_t grid = ctx->num_sms;
    const int64_t want = (sp.n8 + 511) / 512;
    if (grid > want) grid = want;
    if (grid > B2Q_RES_THREADS) grid = B2Q_RES_THREADS;   // the mean path reads one partial per thread
    if (grid < 1) grid = 1;
    b2q_timed_launch tl(ctx, B2Q_KIND_FUSED_FWD, 12.0 * (double)n, st);
    // shared memory sized to the chunk (a small tensor must not force the SMs to the maximum carve-out)
    const int64_t words_per_cta = (sp.n8 + grid - 1) / grid;
    int smem_words = B2Q_RES_SMEM_BYTES / 32;
    if (words_per_cta < smem_words) smem_words = (int)((words_per_cta + 3) & ~(int64_t)3);
    const size_t smem_bytes = (size_t)smem_words * 32;
    if (clip_mode == B2Q_CLIP_SYM)
        b2q_launch_smem(ctx, B2Q_RES_KERNEL(B2Q_CLIP_SYM), (unsigned)grid, B2Q_RES_THREADS, smem_bytes, st, x, y,
                        sp, slot, u, qlevel, ctx->fast_div, clip_with_fresh, (float)n, smem_words);
    else
        b2q_launch_smem(ctx, B2Q_RES_KERNEL(B2Q_CLIP_NONE), (unsigned)grid, B2Q_RES_THREADS, smem_bytes, st, x, y,
                        sp, slot, u, qlevel, ctx->fast_div, clip_with_fresh, (float)n, smem_words);
#undef B2Q_RES_KERNEL
    B2Q_LAUNCH_CHECK(ctx);
    *done = 1;
    return 0;
}
