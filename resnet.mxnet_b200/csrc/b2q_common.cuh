// Shared device/host infrastructure of libb2q.so (sm_100a only).
#pragma once
#include <utility>
#include <cuda_runtime.h>
#include <stdint.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b2q.h"

#define B2Q_MAX_PIECES 16384   // per-launch partial results kept in a workspace slot
#define B2Q_MAX_GROUPS 8192    // thresholds per tensor (channels / groups)
#define B2Q_NSLOTS 64          // reduction workspaces per ctx
#define B2Q_NRINGS 4           // ... in per-stream rings of B2Q_NSLOTS / B2Q_NRINGS
#define B2Q_THREADS 256

// One in-flight reduction.  `ticket` counts finished blocks; the last block combines `partial` in a fixed
// order, applies the threshold update and resets the ticket, so a slot is reusable without a memset.
struct b2q_slot {
    unsigned int ticket;
    unsigned int epoch;           // device-side use counter of max64 (advanced by the consumer kernel, so a CUDA graph
                                  // that replays the pair keeps producing fresh tags)
    unsigned long long max64;     // (epoch << 32 | float bits) for the deferred max reduction: atomicMax, never reset
    float scale[B2Q_MAX_GROUPS];  // threshold T the following QDQ kernel scales with (when it is not aux)
    float clip[B2Q_MAX_GROUPS];   // threshold the following QDQ kernel clips with (when it differs)
    double partial[B2Q_MAX_PIECES];
    unsigned int row_ticket[B2Q_MAX_GROUPS];   // per-channel "last block finishes" tickets (bnstat_fold_kernel); self-resetting
};

// Optional per-kernel timing (option "timing"): CUDA events around every launch of the flat kernels, read
// back with b2q_timing_read().  bench.py uses it for the roofline numbers; it is off by default.
#define B2Q_KIND_REDUCE_FLAT 1
#define B2Q_KIND_QDQ_HOT 2
#define B2Q_KIND_BWD_STE 3
#define B2Q_KIND_BWD_MASK 4
#define B2Q_KIND_OTHER 5
#define B2Q_KIND_FUSED_FWD 6   // single-launch forward of a resident tensor (12 B/element algorithmic)
#define B2Q_NKINDS 7

struct b2q_timing_rec {
    int kind;
    double bytes;
    cudaEvent_t e0, e1;
};

struct b2q_ctx {
    int device = 0;
    int num_sms = 0;
    b2q_slot* slots = nullptr;      // device
    std::mutex mu;                  // guards the ring table, the timing records and the event pool
    bool ring_used[B2Q_NRINGS] = {false, false, false, false};
    cudaStream_t ring_stream[B2Q_NRINGS] = {nullptr, nullptr, nullptr, nullptr};
    unsigned int ring_next[B2Q_NRINGS] = {0, 0, 0, 0};
    int shared_rings = 0;           // times a fifth stream had to share ring 0 (option "shared_slot_rings", read-only)
    long long launches = 0;
    // run-time knobs (never change results)
    int blocks_per_sm = 4096;        // flat QDQ / backward sweeps: grid = min(tiles, SMs x this); the sweep shows one
                                     // 16 KB tile per block (no grid-stride loop) is fastest: dynamic balance over SMs
    int reduce_blocks_per_sm = 4;    // flat reductions that finalise in their last block (partials are re-read)
    int reduce_deferred_blocks_per_sm = 64;   // flat reductions with deferred update (one atomicMax per block)
    int pdl = 1;                              // programmatic dependent launch between consecutive whole-tensor kernels
    int peer_reduce_blocks_per_sm = 8;        // max reductions of the peer-memory exchange (atomicMax + ticket per block)
    int deferred = 1;                // consumer-side threshold update in the fused whole-tensor forward
    int reverse = 1;                 // QDQ sweep walks descending addresses when the tensor exceeds reverse_min_mb MiB
    int reverse_min_mb = 96;            // (option) tensors above this many MiB are swept in descending order
    int fast_div = 1;
    int dorefa_tanh_max = 0;         // 1: DoReFa takes max|tanh(w)| element-wise instead of tanhf(max|w|) (same float)
    int host_ste_copy = 1;           // host-buffer straight-through backward: copy host to host, no PCIe round trip
    int resident = 0;                // single-launch forward for tensors that fit on chip (shared memory + L2).  OFF by
                                     // default: on B200 the 126 MB L2 already serves the sweep of a <= 51 MB tensor (ncu in
                                     // step: 3.6 MB of DRAM reads for a 51.4 MB tensor), so the kernel saves no traffic,
                                     // and its barrier latency chain costs more than two PDL-chained launches
                                     // (profiles/r02b_resident_ab.md)
    int resident_max_mb = 72;        // largest tensor (MB) that takes the single-launch resident forward
    int peer_mode = 1;               // 1: ticket-free reduction, the sweep's first block publishes to the peers; 0: r1 kernels;
                                     // 2 / 3: as 1 with one / two further tiles per block staged in shared memory during the wait
    int cluster_fwd = 1;             // small whole-tensor forwards in ONE launch of one thread-block cluster (b2q_cluster.cuh)
    int cluster_max_elems = 0;       // ... up to this many elements for the max-based operators (kernel limit 327 680;
                                     // 0 = off: measured neutral in the step, profiles/r02s_cluster_sweep.log)
    int cluster_max_elems_mean = 147456;   // ... and for the mean-based ones
    int cluster_words_per_cta = 1024;   // grow the cluster (1/2/4/8 CTAs) while a CTA would hold more 256-bit words than this
    int seg_masked = 1;              // batch statistics over rows that are not a multiple of 4 floats: masked aligned 256-bit words
    int stream_reduce = 0;           // 1: segmented / batch-statistics reductions through the TMA-staged ring when eligible (measured slower)
    int stream_stages = 4;           // 16 KB stages per block of that ring
    int stream_icvt = 0;             // float -> double conversions of the statistics kernels: 0 XU pipe, 1 integer pipe, 2 half / half
    int bn_variant = 1;              // bnstat_fold_kernel tuning variant (b2q_bnfold.cu)
    int bn_pieces_per_sm = 8;        // pieces (blocks) per SM the batch-statistics launch is split into
    int peer_allreduce_blocks_per_sm = 2;   // grid cap of peer_allreduce_kernel
    int peer_allreduce_blocks = 148;        // > 0: absolute grid cap (96..148 measured best at 2 GPUs); 0: blocks_per_sm x SMs
    int peer_publish_blocks_per_sm = 16;   // peer_mode 4: grid of the publishing reduction (one ticket per block)
    int peer_stage_early = 0;        // 1: issue the staging copies before the dependency wait instead of right after it
    int peer_timeout_ms = 600000;    // how long a sweep waits for a peer's statistic before it gives up (NaN output + flag)
    int timing = 0;
    std::vector<b2q_timing_rec> recs;
    std::vector<cudaEvent_t> event_pool;
    bool resident_optin[4] = {false, false, false, false};   // > 48 KB dynamic shared memory for fused_resident_kernel<..>
    bool rows_cta_optin[2] = {false, false};   // > 48 KB dynamic shared memory enabled for rows_cta_kernel<max / sum>
    void* host_state = nullptr;     // staging buffers + streams of the host-buffer entry points (b2q_host.cu)
};

struct b2q_timed_launch {
    b2q_ctx* ctx;
    cudaStream_t st;
    cudaEvent_t e1;
    bool on;
    b2q_timed_launch(b2q_ctx* c, int kind, double bytes, cudaStream_t s) : ctx(c), st(s), e1(nullptr), on(c->timing != 0) {
        if (!on) return;
        std::lock_guard<std::mutex> lk(c->mu);
        cudaEvent_t ev[2];
        for (int i = 0; i < 2; ++i) {
            if (!c->event_pool.empty()) { ev[i] = c->event_pool.back(); c->event_pool.pop_back(); }
            else cudaEventCreate(&ev[i]);
        }
        cudaEventRecord(ev[0], s);
        e1 = ev[1];
        c->recs.push_back({kind, bytes, ev[0], ev[1]});
    }
    ~b2q_timed_launch() { if (on) cudaEventRecord(e1, st); }
};

void b2q_set_error(const std::string& msg);
int b2q_host_release(b2q_ctx* ctx);  // b2q_host.cu

#define B2Q_CHECK_CUDA(expr)                                                                       \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            b2q_set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" +       \
                          __FILE__ + ":" + std::to_string(__LINE__) + ")");                        \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

#define B2Q_REQUIRE(cond, msg)                                                                     \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            b2q_set_error(std::string("b2q: ") + (msg) + " [" #cond "]");                          \
            return 2;                                                                              \
        }                                                                                          \
    } while (0)

#define B2Q_LAUNCH_CHECK(ctx)                                                                      \
    do {                                                                                           \
        (ctx)->launches++;                                                                         \
        B2Q_CHECK_CUDA(cudaGetLastError());                                                        \
    } while (0)

// Launch with the programmatic-stream-serialization attribute: the blocks of this kernel may become resident while the
// previous kernel of the stream drains its last wave (every kernel launched this way starts with b2q_pdl_sync(), which
// waits for the predecessor to complete and flush before touching memory, then lets its own successor do the same).
// It hides the launch latency and prologue of ~300 back-to-back kernels per step; ordering and results are unchanged.
template <typename... KArgs, typename... Args>
static inline void b2q_launch_smem(b2q_ctx* ctx, void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem,
                                   cudaStream_t st, const Args&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr = {};
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = ctx->pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
    if (e == cudaErrorNotSupported && cfg.numAttrs) {
        (void)cudaGetLastError();
        ctx->pdl = 0;
        cfg.numAttrs = 0;
        cudaLaunchKernelEx(&cfg, kernel, args...);
    }
}

template <typename... KArgs, typename... Args>
static inline void b2q_launch(b2q_ctx* ctx, void (*kernel)(KArgs...), unsigned grid, unsigned block,
                              cudaStream_t st, const Args&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(block, 1, 1);
    cfg.stream = st;
    cudaLaunchAttribute attr = {};
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = ctx->pdl ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
    if (e == cudaErrorNotSupported && cfg.numAttrs) {   // a driver without dependent launches: plain launches from now on
        (void)cudaGetLastError();
        ctx->pdl = 0;
        cfg.numAttrs = 0;
        cudaLaunchKernelEx(&cfg, kernel, args...);
    }   // any other error is picked up by B2Q_LAUNCH_CHECK
}

// Reduction workspaces are handed out per STREAM: the 64 slots form four rings of 16, and each of the first four
// distinct streams that use a context gets a ring of its own, so work on different streams (a framework's side stream,
// the host-buffer path's compute stream, a captured graph replayed next to eager calls) never shares a slot.  Within
// one stream a slot is reused 16 calls later, which stream order makes safe.  A fifth stream shares ring 0: streams
// that share a ring must not run concurrently (b2q_get_option(ctx, "shared_slot_rings") counts how often that fallback
// was taken).  The table is guarded by a mutex: MXNet invokes CustomOps from worker threads.
static inline b2q_slot* b2q_take_slot(b2q_ctx* ctx, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    int ring = -1;
    for (int i = 0; i < B2Q_NRINGS; ++i) {
        if (ctx->ring_used[i] && ctx->ring_stream[i] == st) { ring = i; break; }
    }
    if (ring < 0) {
        for (int i = 0; i < B2Q_NRINGS; ++i) {
            if (!ctx->ring_used[i]) { ctx->ring_used[i] = true; ctx->ring_stream[i] = st; ring = i; break; }
        }
    }
    if (ring < 0) { ring = 0; ctx->shared_rings++; }
    const unsigned int i = ctx->ring_next[ring]++;
    return ctx->slots + ring * (B2Q_NSLOTS / B2Q_NRINGS) + (i % (B2Q_NSLOTS / B2Q_NRINGS));
}

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
// 256-bit global accesses (sm_100 LDG.256 / STG.256) with explicit L1/L2 policies.
// Loads  POL 0: plain                                  (reduction pass: lines should stay in L2)
//            1: read-only path, no L1 allocation
//            2: as 1, L2 evict_first                   (last use of the line: the QDQ / backward sweeps)
//            3: no L1 allocation, L2 evict_last        (reduction pass, keep the tail for the QDQ sweep)
// Stores POL 0: plain   1: no L1 allocation, L2 evict_first (pure streaming)   2: L2 evict_last
struct f8 {
    float v[8];
};

#define B2Q_LD8(Q)                                                                                       \
    asm volatile("ld.global" Q ".v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"                                \
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]),   \
                   "=f"(r.v[6]), "=f"(r.v[7])                                                            \
                 : "l"(p))
#define B2Q_ST8(Q)                                                                                       \
    asm volatile("st.global" Q ".v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r.v[0]),          \
                 "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]), "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7]) \
                 : "memory")

template <int POL>
__device__ __forceinline__ f8 ld_f8(const float* p) {
    f8 r;
    if (POL == 0) B2Q_LD8("");
    else if (POL == 1) B2Q_LD8(".nc.L1::no_allocate.L2::evict_normal");
    else if (POL == 2) B2Q_LD8(".nc.L1::no_allocate.L2::evict_first");
    else B2Q_LD8(".L1::no_allocate.L2::evict_last");
    return r;
}

template <int POL>
__device__ __forceinline__ void st_f8(float* p, const f8& r) {
    if (POL == 0) B2Q_ST8("");
    else if (POL == 1) B2Q_ST8(".L1::no_allocate.L2::evict_first");
    else B2Q_ST8(".L2::evict_last");
}

__device__ __forceinline__ void st_i8(int32_t* p, const int (&c)[8]) {
    asm volatile("st.global.v8.s32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(c[0]), "r"(c[1]), "r"(c[2]),
                 "r"(c[3]), "r"(c[4]), "r"(c[5]), "r"(c[6]), "r"(c[7])
                 : "memory");
}

// ---- exact float32 -> float64 on the integer pipe (option stream_icvt; OFF by default) ---------------------------------
// Every sum in this library is accumulated in double (deterministic, correctly rounded results), one conversion per
// element.  F2F.F64.F32 issues on the XU pipe, which ncu shows 64-77 % busy at 44-49 % of DRAM throughput in the
// batch-statistics kernel.  The same value is available from four integer instructions: the float's exponent and
// mantissa fields dropped into a double's (shift by 3 / 29 bits) give x * 2^-896 EXACTLY for every finite x -- zeros and
// denormals included, because a float denormal m * 2^-149 lands on the double denormal m * 2^-1045 -- and one
// multiplication by 2^896 (exact) restores the scale.  Only Inf / NaN (exponent field 255) do not map: the caller tests
// for them and redoes the word with the real conversion.  b2q_selftest(5) compares both conversions over all 2^32 bit
// patterns.  MEASURED (profiles/r02i_bnstat_pipes.md): equal bits, but 5-8 % slower than the conversion instruction in
// every kernel tried (four ALU instructions + a DMUL per element cost more issue slots than the XU pipe saves), so the
// kernels keep the instruction; the option remains for A/B runs.
__device__ __forceinline__ double b2q_two_p896() { return __hiloint2double(0x77f00000, 0); }   // 2^896
__device__ __forceinline__ double b2q_two_m896() { return __hiloint2double(0x07f00000, 0); }   // 2^-896

__device__ __forceinline__ double f2d_scaled(float x, bool& special) {           // x * 2^-896
    const unsigned f = __float_as_uint(x), a = f & 0x7fffffffu;
    special |= (a >= 0x7f800000u);
    return __hiloint2double((int)((a >> 3) | (f & 0x80000000u)), (int)(f << 29));
}

__device__ __forceinline__ double f2d_abs_scaled(float x, bool& special) {       // |x| * 2^-896
    const unsigned a = __float_as_uint(x) & 0x7fffffffu;
    special |= (a >= 0x7f800000u);
    return __hiloint2double((int)(a >> 3), (int)(a << 29));
}

__device__ __forceinline__ double f2d_cvt(float x) {   // the conversion instruction, not speculated by the compiler
    double d;
    asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(x));
    return d;
}

// First statement of every kernel launched through b2q_launch (no effect for an ordinary launch).
__device__ __forceinline__ void b2q_pdl_sync() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ---- bulk asynchronous copies (TMA engine, no tensor map) + mbarrier completion ------------------------------------------
// cp.async.bulk moves a contiguous, 16-byte aligned range global -> shared without occupying the threads or their
// registers; the mbarrier it signals counts the bytes that have landed.  Used where a tile is read more than once from
// shared memory (weight rows: statistic pass + QDQ pass; the resident single-launch forward).
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_fence_init() {   // make the initialised barrier visible to the async proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}

__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {}
}

__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// same with an L2 eviction-priority hint (createpolicy): x is read for the last time by the sweep that stages it
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

__device__ __forceinline__ void bulk_g2s_hint(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar,
                                              unsigned long long policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}

// Kernels that stage tiles in shared memory: ask once per KERNEL (keyed by its address -- template instantiations share
// a function-pointer type) for the dynamic shared memory it needs and the largest shared-memory carve-out.
static inline cudaError_t b2q_kernel_smem_once(const void* kernel, size_t dyn_smem) {
    static std::mutex mu;
    static std::map<const void*, size_t> done;
    std::lock_guard<std::mutex> lk(mu);
    auto it = done.find(kernel);
    if (it != done.end() && it->second >= dyn_smem) return cudaSuccess;
    cudaError_t e = cudaSuccess;
    if (dyn_smem > 0) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn_smem);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    done[kernel] = dyn_smem;
    return cudaSuccess;
}

template <typename K>
static inline void b2q_prefer_shared(K kernel) { (void)b2q_kernel_smem_once((const void*)kernel, 0); }

// "Last block finishes" ticket: a release atomic (+ an acquire fence in the one block that draws `last`) instead of
// __threadfence() + atomicAdd.  __threadfence() is a sequentially consistent fence that also invalidates the SM's whole
// L1 (MEMBAR.SC.GPU + CCTL.IVALL) in every block; the release orders the block's earlier writes (partials, atomicMax)
// before the ticket, and the acquire lets the last block (and, through the following barrier, its other threads) see
// everybody's.
__device__ __forceinline__ unsigned int b2q_take_ticket(unsigned int* ticket, unsigned int last) {
    unsigned int t;
    asm volatile("atom.release.gpu.global.add.u32 %0, [%1], 1;" : "=r"(t) : "l"(ticket) : "memory");
    if (t == last) asm volatile("fence.acq_rel.gpu;" ::: "memory");
    return t;
}

// max / min that PROPAGATE NaN (PTX max.NaN / min.NaN, one FMNMX like fmaxf / fminf, which drop it).  The reference's
// comparison-based mx.nd.clip passes a NaN input through and max(|x|) of a tensor holding a NaN is NaN (NumPy oracle
// and the torch-backed shim agree): a diverged network must produce NaN, not be silently quantised.  On the fast QDQ
// path a NaN that survives the clip fails the safety test, so the word is redone with the reference arithmetic.
__device__ __forceinline__ float fmax_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

__device__ __forceinline__ float fmin_nan(float a, float b) {
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax_nan(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide reduce (fixed tree => deterministic); result valid in thread 0.
template <bool IS_MAX>
__device__ __forceinline__ double block_reduce(double v, double* smem /* >= 32 doubles */) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (IS_MAX) v = (double)warp_max((float)v); else v = warp_sum(v);
    if (lane == 0) smem[wid] = v;
    __syncthreads();
    if (wid == 0) {
        double w = (lane < nw) ? smem[lane] : 0.0;
        if (IS_MAX) w = (double)warp_max((float)w); else w = warp_sum(w);
        v = w;
    }
    __syncthreads();
    return v;
}

// mx.nd.sign: sign(0) = 0
__device__ __forceinline__ float mx_sign(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

// mx.nd.clip: comparisons, NaN passes through (src/operator/tensor/matrix_op-inl.h [upstream])
__device__ __forceinline__ float mx_clip(float x, float lo, float hi) {
    return x > hi ? hi : (x < lo ? lo : x);
}

// ------------------------------------------------------------------------------------------------
// Threshold update, shared by the reductions' last block and the stand-alone kernel.
// Every step is a separately rounded float32 operation, as in the reference's chain of mx.nd calls.
// ------------------------------------------------------------------------------------------------
struct UpdateArgs {
    int mode;        // B2Q_UPD_* or 0 (only emit stat)
    int write_aux;   // aux is updated (is_train); otherwise aux is left alone
    int use_aux_as_scale;  // the QDQ pass scales with aux (after update) rather than with the fresh statistic
    float p0, p1;    // (ema_decay, 1-ema_decay) or (ktimes, lamda)
    float* aux;      // [groups]
    float* scale_out;  // [groups] threshold to scale with
    float* clip_out;   // [groups] threshold to clip with (may be null)
    float* stat_out;   // [groups] raw statistic (may be null)
    unsigned int* seq_counter;   // peer exchange: call counter advanced once per reduction (deferred kernels only)
};

// Pure part of the update: from the old aux value and the reduced statistic to (batch threshold, new aux).
__device__ __forceinline__ void compute_update(int mode, float p0, float p1, float a, float stat, float& fresh,
                                               float& next) {
    fresh = stat;
    next = a;
    switch (mode) {
        case B2Q_UPD_STORE: next = stat; break;
        case B2Q_UPD_EMA: next = __fadd_rn(__fmul_rn(a, p0), __fmul_rn(stat, p1)); break;
        case B2Q_UPD_GDRQ_WEIGHT: fresh = __fmul_rn(p0, stat); next = fresh; break;
        case B2Q_UPD_GDRQ_ACT:
            fresh = __fmul_rn(p0, stat);
            next = __fadd_rn(a, __fmul_rn(p1, __fsub_rn(a, fresh)));
            break;
        case B2Q_UPD_TWICE_STORE: fresh = __fmul_rn(2.f, stat); next = fresh; break;
        case B2Q_UPD_TWICE_EMA:
            fresh = __fmul_rn(2.f, stat);
            next = __fadd_rn(__fmul_rn(a, p0), __fmul_rn(fresh, p1));
            break;
        default: break;
    }
}

__device__ __forceinline__ void apply_update(const UpdateArgs& u, int g, float stat) {
    if (u.stat_out) u.stat_out[g] = stat;
    if (u.mode == 0) return;
    const float a = u.aux ? u.aux[g] : 0.f;
    float fresh, next;
    compute_update(u.mode, u.p0, u.p1, a, stat, fresh, next);
    if (u.write_aux && u.aux) u.aux[g] = next;
    const float after = u.write_aux ? next : a;
    if (u.scale_out) u.scale_out[g] = u.use_aux_as_scale ? after : fresh;
    if (u.clip_out) u.clip_out[g] = fresh;
}

// Deferred ("consumer-side") threshold update for the fused whole-tensor forward: the reduction kernel publishes only
// its result (max: one atomicMax per block on an epoch-tagged 64-bit word, exact and order independent, no reset
// needed because a newer epoch always wins; sums: per-block double partials combined by every consumer block in the
// same fixed order) plus a snapshot of the old aux value.  The QDQ sweep that follows derives the thresholds in
// registers and its block 0 writes the new aux.  This removes the serialized fence -> ticket -> last-block -> update
// chain (~3 us per node, profiles/r01b_sweep.csv) from between the two kernels.
struct DeferredUpdate {
    const double* partial;   // sums: [n_partials] written by reduce_flat_kernel<false, .., FINALIZE=false>
    const unsigned long long* max64;   // max: epoch-tagged atomicMax word written by reduce_flat_kernel<true, ..>
    unsigned int* epoch;     // slot->epoch: the sweep's block 0 stores the consumed tag here
    const float* aux_old;    // snapshot of aux[0] taken by the reduction kernel
    int n_partials;
    int is_max;
    float count;             // elements (for means)
    UpdateArgs u;            // mode / write_aux / use_aux_as_scale / p0 / p1 / aux (destination)
};
