// Single-launch forward for SMALL whole-tensor nodes (by default the mean-based operators -- GDRQ_PY, GDRQ_Fold_BN data --
// up to 144 K elements; options cluster_max_elems / cluster_max_elems_mean; the kernel itself handles up to 320 K): one
// thread-block CLUSTER of 1 / 2 / 4 / 8 CTAs.
//
// A weight node is ~1 % of a step's bytes but, as reduction + sweep, two launches of ~4 us each -- 54 (ResNet-50) to
// 105 (ResNeXt-101) times per step in the one-call-per-node mode a CustomOp framework drives.  Here the tensor is read
// from HBM ONCE into the shared memory of the cluster's CTAs while the statistic is taken (max|x| or the fp64 sum of
// |x|); the per-CTA partials cross the cluster through distributed shared memory (mapa + ld.shared::cluster) between
// two hardware cluster barriers (barrier.cluster.arrive.release / wait.acquire) -- no global workspace, no atomics, no
// ticket, no second launch; every CTA combines the partials in rank order (so all hold the same bits), derives the
// threshold in registers exactly as the two-kernel path does (compute_update), and quantises its part out of shared
// memory.  HBM traffic 8 B/element instead of 12, one launch instead of two.  Bit-identical to the two-kernel path for
// the max-based operators; the mean-based ones differ only in how the (exactly accumulated, double) sum is
// partitioned, as the other reduction paths do among themselves.
//
// Measured (tools/cluster_sweep.py, profiles/r02s_cluster_sweep.log; device time per node from a CUDA graph of 64 nodes):
// 3.0-3.8 us against 3.5-3.9 us for two launches up to 36 K elements (max-based; mean-based 3.1-4.3 against 3.6-5.7 up to
// 64 K), equal at 64 K, and SLOWER beyond (8.3 against 5.1 us at 256 K elements: eight SMs move 1 MB slower than the
// sixty-odd the two-kernel path spreads it over) -- hence the size limits; on the host side one launch instead of two
// is 9.6 against 11.9 us per call.
//
// Compared with fused_resident_kernel (b2q_resident.cuh, persistent grid + software grid barrier through global memory,
// measured slower than two launches): the barrier is the cluster hardware barrier (~1 us), and the kernel is only used
// where one cluster can hold the tensor.
#pragma once
#include "b2q_common.cuh"
#include "b2q_qdq.cuh"
#include "b2q_reduce.cuh"

#define B2Q_CL_THREADS 512
#define B2Q_CL_MAX_WORDS 5120            // 256-bit words staged per CTA: 160 KB of shared memory
#define B2Q_CL_MAX_CTAS 8                // portable cluster size

__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// the double at `local` (a shared-memory address of THIS CTA) as it is in CTA `rank` of the cluster
__device__ __forceinline__ double ld_dsmem_f64(const double* local, unsigned rank) {
    unsigned remote;
    double v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local)), "r"(rank));
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(remote) : "memory");
    return v;
}

template <bool IS_MAX, int CLIP>
__global__ void __launch_bounds__(B2Q_CL_THREADS, 1)
fused_cluster_kernel(const float* __restrict__ x, float* __restrict__ y, FlatSplit sp, UpdateArgs u, float qlevel, int fast,
                     int clip_with_fresh, float count, int n_ctas) {
    extern __shared__ __align__(128) float s_stage[];
    __shared__ double s_red[32];
    __shared__ double s_part;        // this CTA's partial statistic, read by the whole cluster
    __shared__ float s_bcast[2];     // [0] old aux, [1] combined statistic
    b2q_pdl_sync();
    const int tid = threadIdx.x;
    const unsigned rank = cluster_ctarank();
    const int64_t W = (sp.n8 + n_ctas - 1) / n_ctas;                 // 256-bit words per CTA
    const int64_t w0 = (int64_t)rank * W;
    int64_t nw = sp.n8 - w0;
    if (nw > W) nw = W;
    if (nw < 0) nw = 0;
    const float* xb = x + sp.head + 8 * w0;
    float* yb = y + sp.head + 8 * w0;
    if (tid == 0) s_bcast[0] = u.aux ? u.aux[0] : 0.f;               // before the first cluster barrier: CTA 0 overwrites it after
    // ---- phase 1: one pass over HBM: stage + statistic ----
    double acc = 0.0;
    float mx = 0.f;
    float4* s4 = reinterpret_cast<float4*>(s_stage);
    for (int64_t i0 = tid; i0 < nw; i0 += 2 * B2Q_CL_THREADS) {
        const int64_t i1 = i0 + B2Q_CL_THREADS;
        f8 v0, v1;
        v0 = ld_f8<B2Q_QDQ_LDPOL>(xb + 8 * i0);
        if (i1 < nw) v1 = ld_f8<B2Q_QDQ_LDPOL>(xb + 8 * i1);
        else {
#pragma unroll
            for (int j = 0; j < 8; ++j) v1.v[j] = 0.f;
        }
        s4[2 * i0] = make_float4(v0.v[0], v0.v[1], v0.v[2], v0.v[3]);
        s4[2 * i0 + 1] = make_float4(v0.v[4], v0.v[5], v0.v[6], v0.v[7]);
        if (i1 < nw) {
            s4[2 * i1] = make_float4(v1.v[0], v1.v[1], v1.v[2], v1.v[3]);
            s4[2 * i1 + 1] = make_float4(v1.v[4], v1.v[5], v1.v[6], v1.v[7]);
        }
        acc8<IS_MAX>(acc, mx, v0);
        acc8<IS_MAX>(acc, mx, v1);
    }
    if (rank == 0) {  // the (at most 14) unaligned scalars
        if ((int64_t)tid < sp.head) acc1<IS_MAX>(acc, mx, x[tid]);
        if ((int64_t)tid < sp.tail) acc1<IS_MAX>(acc, mx, x[sp.head + 8 * sp.n8 + tid]);
    }
    const double r = block_reduce<IS_MAX>(IS_MAX ? (double)mx : acc, s_red);
    if (tid == 0) s_part = r;
    // ---- the partials cross the cluster through distributed shared memory ----
    cluster_arrive();
    cluster_wait();
    if (tid == 0) {
        double tot = ld_dsmem_f64(&s_part, 0);
        for (int c = 1; c < n_ctas; ++c) {
            const double p = ld_dsmem_f64(&s_part, (unsigned)c);
            if (IS_MAX) tot = (double)fmax_nan((float)tot, (float)p); else tot += p;
        }
        s_bcast[1] = IS_MAX ? (float)tot : __fdiv_rn((float)tot, count);
    }
    __syncthreads();
    cluster_arrive();      // "I have read everybody's partial": matched by the wait before this CTA exits
    // ---- threshold, in registers (same arithmetic as the two-kernel path) ----
    const float stat = s_bcast[1];
    const float a_old = s_bcast[0];
    float fresh, next;
    compute_update(u.mode, u.p0, u.p1, a_old, stat, fresh, next);
    const float after = u.write_aux ? next : a_old;
    const float T = u.use_aux_as_scale ? after : fresh;
    const float Tc = clip_with_fresh ? fresh : T;
    if (rank == 0 && tid == 0 && u.write_aux && u.aux) u.aux[0] = next;
    const bool needs_pos = (CLIP != B2Q_CLIP_NONE && CLIP != B2Q_CLIP_PACT);
    const QScale s = make_qscale(T, qlevel, fast != 0 && !(needs_pos && !(Tc >= 0.f)));
    // ---- phase 2: quantise-dequantise out of shared memory ----
    for (int64_t i = tid; i < nw; i += B2Q_CL_THREADS) {
        const float4 a = s4[2 * i], b = s4[2 * i + 1];
        f8 in, o;
        in.v[0] = a.x; in.v[1] = a.y; in.v[2] = a.z; in.v[3] = a.w;
        in.v[4] = b.x; in.v[5] = b.y; in.v[6] = b.z; in.v[7] = b.w;
        qdq8<CLIP>(in, o, Tc, s);
        st_f8<0>(yb + 8 * i, o);
    }
    if (rank == 0) {  // unaligned head / tail scalars
        int64_t idx = -1;
        if ((int64_t)tid < sp.head) idx = tid;
        else if ((int64_t)tid - sp.head < sp.tail) idx = sp.head + 8 * sp.n8 + ((int64_t)tid - sp.head);
        if (idx >= 0) y[idx] = __fmul_rn(quant_code_exact(clip_value(CLIP, x[idx], Tc), s.q), s.q);
    }
    cluster_wait();        // nobody's shared memory goes away while a peer may still read its partial
}

// launch of one cluster with the dependent-launch attribute
template <typename... KArgs, typename... Args>
static inline cudaError_t b2q_launch_cluster(b2q_ctx* ctx, void (*kernel)(KArgs...), unsigned n_ctas, unsigned block, size_t smem,
                                             cudaStream_t st, const Args&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_ctas, 1, 1);
    cfg.blockDim = dim3(block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2] = {};
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = n_ctas;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = ctx->pdl ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// *done = 0: not eligible (caller takes the resident / two-kernel path).
template <bool IS_MAX>
[[maybe_unused]] static int launch_fused_cluster(b2q_ctx* ctx, const float* x, float* y, int64_t n, UpdateArgs u, float qlevel,
                                                 int clip_mode, int clip_with_fresh, cudaStream_t st, int* done) {
    *done = 0;
    // measured crossover against two PDL-chained launches (profiles/r02s_cluster_sweep.log): 64 K elements for the max-based
    // operators, 144 K for the mean-based ones (whose two-kernel path ends in a serial last-block combination).  In a whole
    // step the max-based gain (0.3 us per small node) disappears, so only the mean-based operators use it by default
    // (cluster_max_elems = 0, cluster_max_elems_mean = 147 456): MobileNet-v1 GDRQ +0.4 %, ResNeXt-101 neutral.
    if (!ctx->cluster_fwd || n > (int64_t)(IS_MAX ? ctx->cluster_max_elems : ctx->cluster_max_elems_mean)) return 0;
    if (!(clip_mode == B2Q_CLIP_NONE || clip_mode == B2Q_CLIP_SYM) || u.stat_out != nullptr || u.scale_out != nullptr ||
        u.clip_out != nullptr) return 0;
    FlatSplit sp = b2q_flat_split(x, n);
    if (!same_misalignment(x, y) || sp.head > B2Q_THREADS || sp.n8 < 1) return 0;
    if (sp.n8 > (int64_t)B2Q_CL_MAX_CTAS * B2Q_CL_MAX_WORDS) return 0;
    // 1 / 2 / 4 / 8 CTAs: at most `cluster_words_per_cta` words each while the cluster can still grow
    int n_ctas = 1;
    const int64_t per = ctx->cluster_words_per_cta > 0 ? ctx->cluster_words_per_cta : 1024;
    while (n_ctas < B2Q_CL_MAX_CTAS && (sp.n8 + n_ctas - 1) / n_ctas > per) n_ctas *= 2;
    const int64_t words = (sp.n8 + n_ctas - 1) / n_ctas;
    if (words > B2Q_CL_MAX_WORDS) return 0;
    const size_t smem = (size_t)((words + 3) & ~(int64_t)3) * 32;
    b2q_timed_launch tl(ctx, B2Q_KIND_FUSED_FWD, 12.0 * (double)n, st);
    cudaError_t e;
#define B2Q_CL_GO(C)                                                                                                      \
    do {                                                                                                                  \
        e = b2q_kernel_smem_once((const void*)fused_cluster_kernel<IS_MAX, C>, (size_t)B2Q_CL_MAX_WORDS * 32);             \
        if (e == cudaSuccess)                                                                                             \
            e = b2q_launch_cluster(ctx, fused_cluster_kernel<IS_MAX, C>, (unsigned)n_ctas, B2Q_CL_THREADS, smem, st, x, y, sp, u, \
                                   qlevel, ctx->fast_div, clip_with_fresh, (float)n, n_ctas);                             \
    } while (0)
    if (clip_mode == B2Q_CLIP_SYM) B2Q_CL_GO(B2Q_CLIP_SYM); else B2Q_CL_GO(B2Q_CLIP_NONE);
#undef B2Q_CL_GO
    if (e != cudaSuccess) {   // a platform that refuses the cluster launch: remember it and use the other paths
        (void)cudaGetLastError();
        ctx->cluster_fwd = 0;
        return 0;
    }
    B2Q_LAUNCH_CHECK(ctx);
    *done = 1;
    return 0;
}
