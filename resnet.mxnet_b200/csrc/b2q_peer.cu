// Cross-rank activation-threshold exchange over peer memory (NVLink / NVSwitch), fused into the two kernels of the
// forward pass.  Replaces  [reduce kernel] -> ncclAllReduce(max, 4 bytes) -> [update kernel] -> [QDQ kernel]
// (3 launches + a latency-bound collective on the forward critical path of every activation node) by
//
//   kernel 1  reduce_peer_kernel   local max|x| (one sequence-tagged atomicMax per block: exact, order independent, never
//                                  reset because a newer sequence always wins) or fp64 partial sums of |x|; the block
//                                  that takes the last ticket stores (sequence << 32 | float bits) into slot
//                                  [sequence % 64][rank] of EVERY rank's mailbox with 8-byte system-scope stores
//                                  (P2P over NVLink for peers) and snapshots the old aux.  The stores travel while the
//                                  kernel drains and the sweep launches.
//   kernel 2  the QDQ sweep        every warp reads its OWN (local) mailbox -- one lane per rank, no shared memory, no
//                                  barrier; the first wave waits until all `world` entries of the slot carry this
//                                  sequence -- takes the max over ranks (exact, order independent), applies the EMA /
//                                  first-batch / alpha update in registers and sweeps; block 0 writes the new aux.
//
// Every rank launches the same sequence of calls (data parallel), so rank A's sweep only ever waits for kernels that
// rank B has already enqueued or will enqueue without depending on A's later work: progress is guaranteed as long as
// all ranks run on different GPUs.  Two mailbox slots alternate: no rank can be more than one call ahead because
// each call needs every rank's value.  The poll is bounded: after ~20 s it traps (CUDA error) instead of hanging.
#include <cstring>

#include "b2q_common.cuh"
#include "b2q_qdq.cuh"
#include "b2q_reduce.cuh"

#define B2Q_CTX(ctx)                               \
    B2Q_REQUIRE((ctx) != nullptr, "null context"); \
    B2Q_CHECK_CUDA(cudaSetDevice((ctx)->device))

#define B2Q_PEER_MAX_RANKS 16
#define B2Q_PEER_SLOTS 64
#define B2Q_PEER_BOX_BYTES (B2Q_PEER_SLOTS * B2Q_PEER_MAX_RANKS * 8)
#define B2Q_PEER_MAX_SMS 256
#define B2Q_PEER_STATE_BYTES 256                    // device-side sequence state (so CUDA graphs can replay)
#define B2Q_PEER_BYTES (B2Q_PEER_BOX_BYTES + B2Q_PEER_STATE_BYTES + B2Q_PEER_MAX_SMS * 128)   // + per-SM resolved words

// Device-side state behind the mailbox words (own mailbox + B2Q_PEER_BOX_BYTES); every rank issues the same calls, so
// the sequence numbers stay in step, and a CUDA graph that replays the kernels publishes fresh numbers every time.
struct PeerState {
    unsigned int done;              // calls published so far; advanced by the last block of the reduction, after
                                    // every block of that kernel has read it (they read it before taking a ticket)
    unsigned int seq;               // sequence number of the call in flight, for the sweep
    unsigned long long local_max;   // (seq << 32 | bits): running local max|x| of the call in flight
    unsigned int timeout_seq;       // != 0: a sweep gave up waiting for a peer at this sequence number (its output is NaN)
    unsigned int timeout_rank;      // the first rank whose statistic was missing
    unsigned int ar_seq;            // allreduce calls completed (peer_allreduce_kernel; device-side so graphs replay)
    unsigned int ar_ticket;         // blocks of the allreduce in flight that have finished their slice
    unsigned int ar_timeout_seq;    // != 0: an allreduce gave up waiting for a peer (the buffer is not valid)
    unsigned int ar_timeout_rank;
};

struct PeerBoxes {
    unsigned long long* box[B2Q_PEER_MAX_RANKS];
    int rank, world;
    PeerState* state;
    unsigned long long* resolved;   // [B2Q_PEER_MAX_SMS][16]: per-SM copy of the all-rank statistic, sequence-tagged
};

__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

#define B2Q_PEER_SPIN_LIMIT (1u << 24)   // x >= 64 ns sleeps, then trap: a peer never arrived

// kernel 1: whole-tensor max|x| (IS_MAX) or mean|x| (sum of |x| in fp64, rounded once, / count -- the arithmetic of
// reduce_flat_kernel); the last block publishes to every rank's mailbox
template <bool IS_MAX, int UNROLL, int LDPOL>
__global__ void __launch_bounds__(B2Q_THREADS)
reduce_peer_kernel(const float* __restrict__ x, FlatSplit sp, b2q_slot* slot, const float* aux, PeerBoxes pb,
                   float count) {
    b2q_pdl_sync();
    __shared__ double smem[32];
    __shared__ unsigned int s_ticket;
    __shared__ float s_tot;
    __shared__ unsigned int s_seq;
    PeerState* state = pb.state;
    float mx = 0.f;
    double acc = 0.0;
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
    if (blockIdx.x == 0) {
        if ((int64_t)threadIdx.x < sp.head) acc1<IS_MAX>(acc, mx, x[threadIdx.x]);
        if ((int64_t)threadIdx.x < sp.tail) acc1<IS_MAX>(acc, mx, x[sp.head + 8 * sp.n8 + threadIdx.x]);
    }
    const double r = block_reduce<IS_MAX>(IS_MAX ? (double)mx : acc, smem);
    if (IS_MAX) {
        if (threadIdx.x == 0) {
            const unsigned int seq = __ldcg(&state->done) + 1u;
            atomicMax(&state->local_max, ((unsigned long long)seq << 32) | __float_as_uint((float)r));
            s_ticket = b2q_take_ticket(&slot->ticket, gridDim.x - 1);
            if (s_ticket == gridDim.x - 1) {   // everybody's maximum is in the word: read it back, no second reduction
                s_tot = __uint_as_float((unsigned int)(__ldcg(&state->local_max) & 0xffffffffull));
                s_seq = seq;
            }
        }
        __syncthreads();
        if (s_ticket != gridDim.x - 1) return;
    } else {
        if (threadIdx.x == 0) {
            slot->partial[blockIdx.x] = r;
            s_ticket = b2q_take_ticket(&slot->ticket, gridDim.x - 1);
        }
        __syncthreads();
        if (s_ticket != gridDim.x - 1) return;
        double a = 0.0;   // fixed-order combine of the fp64 partial sums
        for (unsigned int i = threadIdx.x; i < gridDim.x; i += blockDim.x) a += __ldcg(&slot->partial[i]);
        const double tot = block_reduce<false>(a, smem);
        if (threadIdx.x == 0) {
            s_tot = __fdiv_rn((float)tot, count);
            s_seq = __ldcg(&state->done) + 1u;
        }
        __syncthreads();
    }
    if ((int)threadIdx.x < pb.world) {   // one lane per destination rank: 8-byte P2P store, first thing the block does
        const unsigned long long word = ((unsigned long long)s_seq << 32) | __float_as_uint(s_tot);
        st_sys_u64(pb.box[threadIdx.x] + (size_t)(s_seq & 1u) * B2Q_PEER_MAX_RANKS + pb.rank, word);
    }
    if (threadIdx.x == 0) {
        slot->scale[0] = aux[0];   // snapshot of the old threshold for the sweep
        slot->ticket = 0;
        state->seq = s_seq;
        state->done = s_seq;
    }
}

// Prologue of the sweep: the statistic maximised over all ranks.  Tens of thousands of short blocks asking one L2 line
// the same question is what costs (measured: 8 requests per block to the mailbox line added 24 us per node), so the
// answer is cached per SM: the first warps on an SM read the mailbox (one lane per rank; they wait there until all
// `world` entries carry this sequence), take the max (exact, order independent) and store it, sequence-tagged, in that
// SM's own line of `resolved`; every later block on the SM finds it with ordinary L1-cacheable loads -- the cost of the
// single-GPU kernel's prologue.  A stale or missing line only ever sends a warp down the mailbox path again.
__device__ __forceinline__ float peer_gather(const PeerBoxes& pb) {
    const unsigned int seq = pb.state->seq;   // written by this call's reduction kernel (same stream); L1-cacheable
    unsigned int smid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    unsigned long long* mine = pb.resolved + (size_t)(smid % B2Q_PEER_MAX_SMS) * 16;   // 128 bytes apart
    unsigned long long w;
    asm volatile("ld.global.ca.u64 %0, [%1];" : "=l"(w) : "l"(mine) : "memory");
    if ((unsigned int)(w >> 32) == seq) return __uint_as_float((unsigned int)(w & 0xffffffffull));   // warp-uniform
    const int lane = threadIdx.x & 31;
    float v = 0.f;
    if (lane < pb.world) {
        const unsigned long long* p = pb.box[pb.rank] + (size_t)(seq & 1u) * B2Q_PEER_MAX_RANKS + lane;
        unsigned long long m = ld_sys_u64(p);
        unsigned int spins = 0;
        while ((unsigned int)(m >> 32) != seq) {
            __nanosleep(64 + 8 * (threadIdx.x >> 5));
            m = ld_sys_u64(p);
            if (++spins > B2Q_PEER_SPIN_LIMIT) {           // a peer never arrived (~20 s): NaN statistic + flag, no trap
                pb.state->timeout_seq = seq;
                pb.state->timeout_rank = (unsigned int)lane;
                m = ((unsigned long long)seq << 32) | 0x7fc00000ull;
                break;
            }
        }
        v = __uint_as_float((unsigned int)(m & 0xffffffffull));
    }
    v = warp_max(v);
    if (lane == 0) *mine = ((unsigned long long)seq << 32) | __float_as_uint(v);
    return v;
}

// kernel 2: QDQ sweep whose prologue gathers every rank's statistic from the local mailbox
template <int CLIP, int UNROLL, int LDPOL, int STPOL>
__global__ void __launch_bounds__(B2Q_THREADS)
qdq_peer_kernel(const float* __restrict__ x, float* __restrict__ y, FlatSplit sp, PeerBoxes pb, const b2q_slot* slot,
                UpdateArgs u, float qlevel, int fast, int reverse, int clip_with_fresh) {
    const float* xb = x + sp.head;
    float* yb = y + sp.head;
    const int64_t tile = (int64_t)B2Q_THREADS * UNROLL;
    const int64_t ntiles = (sp.n8 + tile - 1) / tile;
    // The first tile's loads go out before the dependency wait and before the mailbox is read: the preceding kernel (this
    // call's reduction) only READS x and has itself waited for x's producer, so the HBM latency overlaps its last wave.
    f8 v[UNROLL];
    int64_t tt = blockIdx.x;
    {
        const int64_t t = reverse ? (ntiles - 1 - tt) : tt;
        const int64_t base = t * tile + threadIdx.x;
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) {
            const int64_t i = base + (int64_t)k * B2Q_THREADS;
            if (tt < ntiles && i < sp.n8) v[k] = ld_f8<LDPOL>(xb + 8 * i);
        }
    }
    b2q_pdl_sync();
    const float a_old = slot->scale[0];
    const float stat = peer_gather(pb);
    float fresh, next;
    compute_update(u.mode, u.p0, u.p1, a_old, stat, fresh, next);
    const float T = next;
    const float Tc = clip_with_fresh ? fresh : T;   // fold_bn_v1_gdrq.py:67 clips with the batch threshold
    if (blockIdx.x == 0 && threadIdx.x == 0) u.aux[0] = next;
    const QScale s = make_qscale(T, qlevel, fast != 0 && !(CLIP != B2Q_CLIP_NONE && !(Tc >= 0.f)));
    for (; tt < ntiles; tt += gridDim.x) {
        const int64_t t = reverse ? (ntiles - 1 - tt) : tt;
        const int64_t base = t * tile + threadIdx.x;
        if (tt != (int64_t)blockIdx.x) {
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
    if (blockIdx.x == 0) {
        const int64_t tid = threadIdx.x;
        int64_t idx = -1;
        if (tid < sp.head) idx = tid;
        else if (tid - sp.head < sp.tail) idx = sp.head + 8 * sp.n8 + (tid - sp.head);
        if (idx >= 0) {
            y[idx] = __fmul_rn(quant_code_exact(clip_value(CLIP, x[idx], Tc), s.q), s.q);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// peer_mode 4: the deferred max reduction with ONE extra atomic per block -- a release ticket after the tagged atomicMax --
// so that the block which finishes last can store this rank's statistic into every rank's mailbox from INSIDE the
// reduction, instead of the sweep's block 0 doing it after the whole grid has drained, the dependent launch has been
// released and the word has been read back (about 1 us of the 5 us the exchange adds per node).  The sweep
// (qdq_peer2_kernel with `published` = 1) then only polls.  max|x| statistics only; the mean-based operators keep mode 1.
// ------------------------------------------------------------------------------------------------------------------
template <int UNROLL, int LDPOL>
__global__ void __launch_bounds__(B2Q_THREADS)
reduce_flat_publish_kernel(const float* __restrict__ x, FlatSplit sp, b2q_slot* slot, UpdateArgs u, PeerBoxes pb) {
    b2q_pdl_sync();
    __shared__ double smem[32];
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
        for (int k = 0; k < UNROLL; ++k) acc8<true>(acc, mx, v[k]);
    }
    if (blockIdx.x == 0) {  // the (at most 14) unaligned scalars
        if ((int64_t)threadIdx.x < sp.head) acc1<true>(acc, mx, x[threadIdx.x]);
        if ((int64_t)threadIdx.x < sp.tail) acc1<true>(acc, mx, x[sp.head + 8 * sp.n8 + threadIdx.x]);
    }
    const double r = block_reduce<true>((double)mx, smem);
    if (threadIdx.x == 0) {
        const unsigned int tag = slot->epoch + 1u;
        atomicMax(&slot->max64, ((unsigned long long)tag << 32) | __float_as_uint((float)r));
        if (blockIdx.x == 0) {
            if (u.aux) slot->scale[0] = u.aux[0];             // snapshot of the old threshold
            *u.seq_counter = *u.seq_counter + 1u;             // this call's sequence number, before block 0's ticket
        }
        if (b2q_take_ticket(&slot->ticket, gridDim.x - 1) == gridDim.x - 1) {
            // every block's atomicMax and block 0's counter update are visible (release tickets, acquire here)
            slot->ticket = 0;
            const unsigned long long word = *((volatile unsigned long long*)&slot->max64);
            const unsigned int seq = *((volatile unsigned int*)u.seq_counter);
            const unsigned long long out = ((unsigned long long)seq << 32) | (word & 0xffffffffull);
            for (int rk = 0; rk < pb.world; ++rk)
                st_sys_u64(pb.box[rk] + (size_t)(seq & 1u) * B2Q_PEER_MAX_RANKS + pb.rank, out);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// peer_mode 1 (default): no ticket chain at all.
//   kernel 1  the ordinary deferred reduction of the single-GPU forward (reduce_flat_kernel<.., FINALIZE=false>: one tagged
//             atomicMax per block, or fp64 partial sums; block 0 snapshots the old aux and advances the call counter).
//   kernel 2  qdq_peer2_kernel: its blocks are resident behind the reduction (programmatic dependent launch) with their
//             first tile already requested.  Block 0 reads the reduced statistic (max: the tagged word; mean: combines
//             the partials in a fixed order) and stores (sequence << 32 | bits) into every rank's mailbox -- one lane
//             per destination, 8-byte system-scope stores over NVLink.  In every block warp 0 alone looks the statistic up
//             (the SM's cached line, else the mailbox: one lane per rank) and hands it to the block through shared
//             memory, so the mailbox line sees one poller per block instead of one per warp.
// The wait is bounded by TIME (option peer_timeout_ms, default 10 minutes, %globaltimer): on expiry the sweep records
// (sequence, missing rank) in the mailbox state, uses NaN as the statistic -- the output and the threshold turn NaN,
// loudly -- and the context stays alive; b2q_peer_status() reports it to the host.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ float peer_gather_warp0(const PeerBoxes& pb, unsigned int seq, unsigned long long timeout_ns) {
    // called by warp 0 of a block (all 32 lanes)
    unsigned int smid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    unsigned long long* mine = pb.resolved + (size_t)(smid % B2Q_PEER_MAX_SMS) * 16;   // 128 bytes apart
    unsigned long long w;
    asm volatile("ld.global.ca.u64 %0, [%1];" : "=l"(w) : "l"(mine) : "memory");
    if ((unsigned int)(w >> 32) == seq) return __uint_as_float((unsigned int)(w & 0xffffffffull));   // warp-uniform
    const int lane = threadIdx.x & 31;
    float v = 0.f;
    bool late = false;
    if (lane < pb.world) {
        const unsigned long long* p = pb.box[pb.rank] + (size_t)(seq & 1u) * B2Q_PEER_MAX_RANKS + lane;
        unsigned long long m = ld_sys_u64(p);
        if ((unsigned int)(m >> 32) != seq) {
            // The common wait is 2-5 us (NVLink latency + the skew between two ranks' reductions): poll tightly for the
            // first ~8 us (one warp per CTA polls, so the mailbox line is not swamped), then back off.
            const unsigned long long t0 = global_ns();
            unsigned int polls = 0;
            while (true) {
                if (polls < 256) __nanosleep(20);
                else __nanosleep(polls < 4096 ? 200 : 2000);
                m = ld_sys_u64(p);
                if ((unsigned int)(m >> 32) == seq) break;
                if ((++polls & 63u) == 0 && global_ns() - t0 > timeout_ns) { late = true; break; }
            }
        }
        v = late ? __int_as_float(0x7fc00000) : __uint_as_float((unsigned int)(m & 0xffffffffull));
    }
    const unsigned int late_mask = __ballot_sync(0xffffffffu, late);
    v = warp_max(v);                                           // NaN-propagating
    if (late_mask) {
        if (lane == 0) {
            pb.state->timeout_seq = seq;
            pb.state->timeout_rank = (unsigned int)(__ffs(late_mask) - 1);
        }
        return v;   // NaN; not cached, so that every block reports the same
    }
    if (lane == 0) *mine = ((unsigned long long)seq << 32) | __float_as_uint(v);
    return v;
}

template <int CLIP, int UNROLL, int LDPOL, int STPOL>
__global__ void __launch_bounds__(B2Q_THREADS)
qdq_peer2_kernel(const float* __restrict__ x, float* __restrict__ y, FlatSplit sp, PeerBoxes pb, b2q_slot* slot,
                 UpdateArgs u, float qlevel, int fast, int reverse, int clip_with_fresh, int is_max, int n_partials,
                 float count, unsigned long long timeout_ns, int published) {
    __shared__ float s_stat;
    __shared__ double s_red[32];
    const float* xb = x + sp.head;
    float* yb = y + sp.head;
    const int64_t tile = (int64_t)B2Q_THREADS * UNROLL;
    const int64_t ntiles = (sp.n8 + tile - 1) / tile;
    f8 v[UNROLL];
    int64_t tt = blockIdx.x;
    {   // first tile: requested before the dependency wait (the reduction in front of us only reads x)
        const int64_t t = reverse ? (ntiles - 1 - tt) : tt;
        const int64_t base = t * tile + threadIdx.x;
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) {
            const int64_t i = base + (int64_t)k * B2Q_THREADS;
            if (tt < ntiles && i < sp.n8) v[k] = ld_f8<LDPOL>(xb + 8 * i);
        }
    }
    b2q_pdl_sync();
    const unsigned int seq = pb.state->seq;   // advanced by this call's reduction (block 0), stable during this kernel
    if (blockIdx.x == 0) {   // publish this rank's statistic to every rank (own mailbox included)
        float mine;
        if (is_max) {
            const unsigned long long word = slot->max64;
            mine = __uint_as_float((unsigned int)(word & 0xffffffffull));
            if (threadIdx.x == 0) slot->epoch = (unsigned int)(word >> 32);   // consume the tag
        } else {
            double acc = 0.0;
            for (int i = threadIdx.x; i < n_partials; i += blockDim.x) acc += slot->partial[i];
            const double tot = block_reduce<false>(acc, s_red);
            if (threadIdx.x == 0) s_stat = __fdiv_rn((float)tot, count);
            __syncthreads();
            mine = s_stat;
            __syncthreads();
        }
        if (!published && (int)threadIdx.x < pb.world) {   // peer_mode 4: the reduction's last block has done this already
            const unsigned long long word = ((unsigned long long)seq << 32) | __float_as_uint(mine);
            st_sys_u64(pb.box[threadIdx.x] + (size_t)(seq & 1u) * B2Q_PEER_MAX_RANKS + pb.rank, word);
        }
        if (threadIdx.x == 0) pb.state->done = seq;   // keeps the r1 kernels' counter (peer_mode 0) in step
    }
    if (threadIdx.x < 32) {
        const float g = peer_gather_warp0(pb, seq, timeout_ns);
        if (threadIdx.x == 0) s_stat = g;
    }
    __syncthreads();
    const float stat = s_stat;
    const float a_old = slot->scale[0];
    float fresh, next;
    compute_update(u.mode, u.p0, u.p1, a_old, stat, fresh, next);
    const float T = next;
    const float Tc = clip_with_fresh ? fresh : T;   // fold_bn_v1_gdrq.py:67 clips with the batch threshold
    if (blockIdx.x == 0 && threadIdx.x == 0) u.aux[0] = next;
    const QScale s = make_qscale(T, qlevel, fast != 0 && !(CLIP != B2Q_CLIP_NONE && !(Tc >= 0.f)));
    for (; tt < ntiles; tt += gridDim.x) {
        const int64_t t = reverse ? (ntiles - 1 - tt) : tt;
        const int64_t base = t * tile + threadIdx.x;
        if (tt != (int64_t)blockIdx.x) {
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
    if (blockIdx.x == 0) {
        const int64_t tid = threadIdx.x;
        int64_t idx = -1;
        if (tid < sp.head) idx = tid;
        else if (tid - sp.head < sp.tail) idx = sp.head + 8 * sp.n8 + (tid - sp.head);
        if (idx >= 0) {
            y[idx] = __fmul_rn(quant_code_exact(clip_value(CLIP, x[idx], Tc), s.q), s.q);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// peer_mode 2 / 3: qdq_peer2_kernel with its waiting time put to work.  Between the end of this rank's reduction and
// the arrival of the slowest rank's statistic (NVLink hop + skew, ~5 us per node) the memory system is idle: the first
// tile of every resident block has long landed in its registers.  Here a block owns 1 + STAGES consecutive tiles: the
// first is requested into registers before the dependency wait as before; the others are fetched into shared memory
// by bulk asynchronous copies (TMA engine, mbarrier completion, L2 evict-first hint) issued by one thread right AFTER
// the dependency wait, i.e. exactly while warp 0 polls the mailbox.  With 8 blocks per SM that is another
// 128 KB (STAGES = 1) per SM of x on its way during the wait -- 19 / 38 MB over the chip, 3 / 6 us of HBM time that
// no longer follows the wait.  The arithmetic is qdq8 on the same 256-bit words: results are bit-identical.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ f8 lds_f8(const float* p) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f8 r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}

template <int CLIP, int UNROLL, int LDPOL, int STPOL, int STAGES>
__global__ void __launch_bounds__(B2Q_THREADS)
qdq_peer3_kernel(const float* __restrict__ x, float* __restrict__ y, FlatSplit sp, PeerBoxes pb, b2q_slot* slot,
                 UpdateArgs u, float qlevel, int fast, int reverse, int clip_with_fresh, int is_max, int n_partials,
                 float count, unsigned long long timeout_ns, int stage_early) {
    __shared__ __align__(128) float s_tile[STAGES][B2Q_THREADS * UNROLL * 8];
    __shared__ unsigned long long s_bar[STAGES];
    __shared__ float s_stat;
    __shared__ double s_red[32];
    const float* xb = x + sp.head;
    float* yb = y + sp.head;
    const int64_t tile = (int64_t)B2Q_THREADS * UNROLL;
    const int64_t ntiles = (sp.n8 + tile - 1) / tile;
    const int64_t lt0 = (int64_t)blockIdx.x * (1 + STAGES);       // this block's tiles: lt0 .. lt0 + STAGES (logical order)
    f8 v[UNROLL];
    {   // first tile: requested before the dependency wait (the reduction in front of us only reads x)
        const int64_t t = reverse ? (ntiles - 1 - lt0) : lt0;
        const int64_t base = t * tile + threadIdx.x;
#pragma unroll
        for (int k = 0; k < UNROLL; ++k) {
            const int64_t i = base + (int64_t)k * B2Q_THREADS;
            if (lt0 < ntiles && i < sp.n8) v[k] = ld_f8<LDPOL>(xb + 8 * i);
        }
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int sgi = 0; sgi < STAGES; ++sgi) mbar_init(&s_bar[sgi], 1);
        mbar_fence_init();
    }
    auto stage = [&]() {   // thread 0 only
        const unsigned long long pol = l2_policy_evict_first();
#pragma unroll
        for (int sgi = 0; sgi < STAGES; ++sgi) {
            const int64_t lt = lt0 + 1 + sgi;
            if (lt < ntiles) {
                const int64_t t = reverse ? (ntiles - 1 - lt) : lt;
                int64_t words = sp.n8 - t * tile;
                if (words > tile) words = tile;
                const unsigned bytes = (unsigned)words * 32u;
                mbar_expect_tx(&s_bar[sgi], bytes);
                bulk_g2s_hint(&s_tile[sgi][0], xb + 8 * t * tile, bytes, &s_bar[sgi], pol);
            }
        }
    };
    if (stage_early && threadIdx.x == 0) stage();
    b2q_pdl_sync();
    if (!stage_early && threadIdx.x == 0) stage();
    const unsigned int seq = pb.state->seq;   // advanced by this call's reduction (block 0), stable during this kernel
    if (blockIdx.x == 0) {   // publish this rank's statistic to every rank (own mailbox included)
        float mine;
        if (is_max) {
            const unsigned long long word = slot->max64;
            mine = __uint_as_float((unsigned int)(word & 0xffffffffull));
            if (threadIdx.x == 0) slot->epoch = (unsigned int)(word >> 32);   // consume the tag
        } else {
            double acc = 0.0;
            for (int i = threadIdx.x; i < n_partials; i += blockDim.x) acc += slot->partial[i];
            const double tot = block_reduce<false>(acc, s_red);
            if (threadIdx.x == 0) s_stat = __fdiv_rn((float)tot, count);
            __syncthreads();
            mine = s_stat;
            __syncthreads();
        }
        if ((int)threadIdx.x < pb.world) {
            const unsigned long long word = ((unsigned long long)seq << 32) | __float_as_uint(mine);
            st_sys_u64(pb.box[threadIdx.x] + (size_t)(seq & 1u) * B2Q_PEER_MAX_RANKS + pb.rank, word);
        }
        if (threadIdx.x == 0) pb.state->done = seq;
    }
    if (threadIdx.x < 32) {
        const float g = peer_gather_warp0(pb, seq, timeout_ns);
        if (threadIdx.x == 0) s_stat = g;
    }
    __syncthreads();   // also publishes the initialised mbarriers to the whole block
    const float stat = s_stat;
    const float a_old = slot->scale[0];
    float fresh, next;
    compute_update(u.mode, u.p0, u.p1, a_old, stat, fresh, next);
    const float T = next;
    const float Tc = clip_with_fresh ? fresh : T;   // fold_bn_v1_gdrq.py:67 clips with the batch threshold
    if (blockIdx.x == 0 && threadIdx.x == 0) u.aux[0] = next;
    const QScale s = make_qscale(T, qlevel, fast != 0 && !(CLIP != B2Q_CLIP_NONE && !(Tc >= 0.f)));
    if (lt0 < ntiles) {
        const int64_t t = reverse ? (ntiles - 1 - lt0) : lt0;
        const int64_t base = t * tile + threadIdx.x;
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
#pragma unroll
    for (int sgi = 0; sgi < STAGES; ++sgi) {
        const int64_t lt = lt0 + 1 + sgi;
        if (lt < ntiles) {
            const int64_t t = reverse ? (ntiles - 1 - lt) : lt;
            const int64_t base = t * tile + threadIdx.x;
            mbar_wait(&s_bar[sgi], 0);
#pragma unroll
            for (int k = 0; k < UNROLL; ++k) {
                const int64_t i = base + (int64_t)k * B2Q_THREADS;
                if (i < sp.n8) {
                    const f8 w = lds_f8(&s_tile[sgi][8 * (k * B2Q_THREADS + (int)threadIdx.x)]);
                    f8 o;
                    qdq8<CLIP>(w, o, Tc, s);
                    st_f8<STPOL>(yb + 8 * i, o);
                }
            }
        }
    }
    if (blockIdx.x == 0) {
        const int64_t tid = threadIdx.x;
        int64_t idx = -1;
        if (tid < sp.head) idx = tid;
        else if (tid - sp.head < sp.tail) idx = sp.head + 8 * sp.n8 + (tid - sp.head);
        if (idx >= 0) {
            y[idx] = __fmul_rn(quant_code_exact(clip_value(CLIP, x[idx], Tc), s.q), s.q);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Allreduce (sum / max) of one float32 buffer per rank over peer memory, one process per GPU: ONE kernel per rank, no
// NCCL, no host synchronisation.  bufs[r] is rank r's buffer as mapped on this device (CUDA IPC).  Rank r owns the r-th
// slice: it reads that slice from every rank (peer loads over NVLink), combines the values in rank order -- so every
// rank ends up with the same bits -- and stores the result into every rank's buffer (peer stores): reduce-scatter and
// all-gather in one pass, (n-1)/n of the buffer in each direction per device like a ring, but in one hop.
// Two flag barriers through the mailboxes (slots 62 / 63 of the box area, 8-byte system-scope words, sequence-tagged):
//   in   block 0 tells every rank "my buffer is complete" (stream order: everything before this kernel has finished);
//        every block waits until all ranks said so before it reads a peer's data
//   out  the block that finishes last (ticket) tells every rank "my slice is stored everywhere" and then waits for the
//        same word from every rank, so the kernel -- and with it this rank's stream -- completes only when the whole
//        buffer is final here and nobody still reads the old contents.
// All ranks must call it in the same order with the same count (data parallel).  The waits are bounded by
// peer_timeout_ms like the threshold exchange (b2q_peer_status reports an expiry; the buffer is then not valid).
// ------------------------------------------------------------------------------------------------------------------
#define B2Q_AR_SLOT_IN 62
#define B2Q_AR_SLOT_OUT 63

struct ArPtrs {
    float* buf[B2Q_PEER_MAX_RANKS];
};

// all 32 lanes of one warp: wait until every rank's word in `slot` of the own mailbox carries `seq`
__device__ __forceinline__ bool ar_wait_all(const PeerBoxes& pb, int slot, unsigned int seq, unsigned long long timeout_ns) {
    const int lane = threadIdx.x & 31;
    bool late = false;
    if (lane < pb.world) {
        const unsigned long long* p = pb.box[pb.rank] + (size_t)slot * B2Q_PEER_MAX_RANKS + lane;
        unsigned long long m = ld_sys_u64(p);
        if ((unsigned int)m != seq) {
            const unsigned long long t0 = global_ns();
            unsigned int polls = 0;
            while (true) {
                __nanosleep(polls < 256 ? 20 : (polls < 4096 ? 200 : 2000));
                m = ld_sys_u64(p);
                if ((unsigned int)m == seq) break;
                if ((++polls & 63u) == 0 && global_ns() - t0 > timeout_ns) { late = true; break; }
            }
        }
    }
    const unsigned int late_mask = __ballot_sync(0xffffffffu, late);
    if (late_mask && lane == 0) {
        pb.state->ar_timeout_seq = seq;
        pb.state->ar_timeout_rank = (unsigned int)(__ffs(late_mask) - 1);
    }
    return late_mask == 0;
}

template <bool IS_MAX>
__global__ void __launch_bounds__(256)
peer_allreduce_kernel(ArPtrs p, PeerBoxes pb, int64_t begin, int64_t end, int vec, float post_scale,
                      unsigned long long timeout_ns) {
    __shared__ unsigned int s_ticket;
    const unsigned int seq = pb.state->ar_seq + 1u;   // read by every block before it takes its ticket (below)
    if (blockIdx.x == 0 && (int)threadIdx.x < pb.world)
        st_sys_u64(pb.box[threadIdx.x] + (size_t)B2Q_AR_SLOT_IN * B2Q_PEER_MAX_RANKS + pb.rank, (unsigned long long)seq);
    if (threadIdx.x < 32) ar_wait_all(pb, B2Q_AR_SLOT_IN, seq, timeout_ns);
    __syncthreads();
    __threadfence_system();   // acquire: the peers' buffers are complete and visible
    const int n_ranks = pb.world;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t sbegin = begin;   // scalar part: everything when the buffers are not 16-byte aligned, else the ragged tail
    if (vec) {
        // four 128-bit words per thread and pass: the loads of one rank are all in flight before the first is used
        const int64_t w_end = end >> 2;
        for (int64_t i0 = (begin >> 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < w_end; i0 += 4 * stride) {
            float4 acc[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t i = i0 + u * stride;
                acc[u] = i < w_end ? __ldcg(reinterpret_cast<const float4*>(p.buf[0]) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            for (int r = 1; r < n_ranks; ++r) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int64_t i = i0 + u * stride;
                    v[u] = i < w_end ? __ldcg(reinterpret_cast<const float4*>(p.buf[r]) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (IS_MAX) {
                        acc[u].x = fmax_nan(acc[u].x, v[u].x); acc[u].y = fmax_nan(acc[u].y, v[u].y);
                        acc[u].z = fmax_nan(acc[u].z, v[u].z); acc[u].w = fmax_nan(acc[u].w, v[u].w);
                    } else {
                        acc[u].x = __fadd_rn(acc[u].x, v[u].x); acc[u].y = __fadd_rn(acc[u].y, v[u].y);
                        acc[u].z = __fadd_rn(acc[u].z, v[u].z); acc[u].w = __fadd_rn(acc[u].w, v[u].w);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t i = i0 + u * stride;
                if (i >= w_end) continue;
                if (!IS_MAX && post_scale != 1.f) {
                    acc[u].x = __fmul_rn(acc[u].x, post_scale); acc[u].y = __fmul_rn(acc[u].y, post_scale);
                    acc[u].z = __fmul_rn(acc[u].z, post_scale); acc[u].w = __fmul_rn(acc[u].w, post_scale);
                }
                for (int r = 0; r < n_ranks; ++r) __stcg(reinterpret_cast<float4*>(p.buf[r]) + i, acc[u]);
            }
        }
        sbegin = (end >> 2) << 2;
        if (sbegin < begin) sbegin = begin;
    }
    for (int64_t i = sbegin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += stride) {
        float acc = __ldcg(p.buf[0] + i);
        for (int r = 1; r < n_ranks; ++r) {
            const float v = __ldcg(p.buf[r] + i);
            acc = IS_MAX ? fmax_nan(acc, v) : __fadd_rn(acc, v);
        }
        if (!IS_MAX && post_scale != 1.f) acc = __fmul_rn(acc, post_scale);
        for (int r = 0; r < n_ranks; ++r) p.buf[r][i] = acc;
    }
    __threadfence_system();   // release: this block's peer stores are visible before its ticket
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = b2q_take_ticket(&pb.state->ar_ticket, gridDim.x - 1);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    // ---- last block: every slice element of this rank is stored everywhere ----
    __threadfence_system();
    if (threadIdx.x == 0) { pb.state->ar_ticket = 0; pb.state->ar_seq = seq; }
    if ((int)threadIdx.x < pb.world)
        st_sys_u64(pb.box[threadIdx.x] + (size_t)B2Q_AR_SLOT_OUT * B2Q_PEER_MAX_RANKS + pb.rank, (unsigned long long)seq);
    if (threadIdx.x < 32) ar_wait_all(pb, B2Q_AR_SLOT_OUT, seq, timeout_ns);
    __threadfence_system();
}

extern "C" {

int b2q_peer_mailbox_bytes(void) { return B2Q_PEER_BYTES; }

int b2q_peer_status(b2q_ctx* ctx, const void* own_mailbox, uint32_t* timeout_sequence, uint32_t* timeout_rank) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(own_mailbox && timeout_sequence, "null argument");
    PeerState st;
    B2Q_CHECK_CUDA(cudaMemcpy(&st, (const char*)own_mailbox + B2Q_PEER_BOX_BYTES, sizeof(st), cudaMemcpyDeviceToHost));
    *timeout_sequence = st.timeout_seq ? st.timeout_seq : st.ar_timeout_seq;
    if (timeout_rank) *timeout_rank = st.timeout_seq ? st.timeout_rank : st.ar_timeout_rank;
    return 0;
}

int b2q_peer_mailbox_create(b2q_ctx* ctx, void** mailbox, void* ipc_handle_out) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(mailbox && ipc_handle_out, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    B2Q_CHECK_CUDA(cudaMalloc(&p, B2Q_PEER_BYTES));
    B2Q_CHECK_CUDA(cudaMemset(p, 0, B2Q_PEER_BYTES));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        b2q_set_error(std::string("cudaIpcGetMemHandle failed: ") + cudaGetErrorString(e));
        return 1;
    }
    memcpy(ipc_handle_out, &h, sizeof(h));
    *mailbox = p;
    return 0;
}

int b2q_peer_mailbox_open(b2q_ctx* ctx, const void* ipc_handle, void** peer_ptr) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(ipc_handle && peer_ptr, "null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    B2Q_CHECK_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int b2q_peer_mailbox_close(b2q_ctx* ctx, void* peer_ptr) {
    B2Q_CTX(ctx);
    if (peer_ptr) B2Q_CHECK_CUDA(cudaIpcCloseMemHandle(peer_ptr));
    return 0;
}

int b2q_peer_mailbox_destroy(b2q_ctx* ctx, void* mailbox) {
    B2Q_CTX(ctx);
    if (mailbox) B2Q_CHECK_CUDA(cudaFree(mailbox));
    return 0;
}

// shared body: [reduce + publish] -> [poll + update + sweep]
static int peer_quant_fwd(b2q_ctx* ctx, bool is_max, int upd_mode, float p0, float p1, int clip_mode, int clip_with_fresh,
                          float qlevel, const float* x, float* y, float* aux, int64_t n, void* const* mailboxes, int rank,
                          int world, cudaStream_t st) {
    B2Q_REQUIRE(x && y && aux && mailboxes && n >= 1, "bad argument");
    B2Q_REQUIRE(world >= 1 && world <= B2Q_PEER_MAX_RANKS && rank >= 0 && rank < world, "bad rank / world");
    FlatSplit sp = b2q_flat_split(x, n);
    B2Q_REQUIRE(same_misalignment(x, y) && sp.head <= B2Q_THREADS, "peer path needs equally aligned float32 buffers");
    PeerBoxes pb;
    memset(&pb, 0, sizeof(pb));
    for (int r = 0; r < world; ++r) {
        B2Q_REQUIRE(mailboxes[r] != nullptr, "null mailbox pointer");
        pb.box[r] = (unsigned long long*)mailboxes[r];
    }
    pb.rank = rank; pb.world = world;
    pb.state = (PeerState*)((char*)mailboxes[rank] + B2Q_PEER_BOX_BYTES);
    static_assert(sizeof(PeerState) <= B2Q_PEER_STATE_BYTES, "mailbox state area");
    pb.resolved = (unsigned long long*)((char*)mailboxes[rank] + B2Q_PEER_BOX_BYTES + B2Q_PEER_STATE_BYTES);
    b2q_slot* slot = b2q_take_slot(ctx, st);
    const int mode = ctx->peer_mode;
    if (mode >= 1) {
        UpdateArgs ur;
        memset(&ur, 0, sizeof(ur));
        ur.aux = aux;                       // the deferred reduction snapshots the old threshold into slot->scale[0]
        ur.seq_counter = &pb.state->seq;
        int np = 0;
        int rc = 0;
        const int published = (mode == 4 && is_max) ? 1 : 0;
        if (published) {   // the reduction's last block publishes (one release ticket per block on top of the atomicMax)
            B2Q_REQUIRE(sp.head <= B2Q_THREADS, "peer path needs equally aligned float32 buffers");
            const int64_t rgrid = b2q_flat_grid(ctx, sp.n8, B2Q_REDUCE_UNROLL, ctx->peer_publish_blocks_per_sm > 0 ? ctx->peer_publish_blocks_per_sm : ctx->reduce_deferred_blocks_per_sm);
            b2q_timed_launch tl(ctx, B2Q_KIND_REDUCE_FLAT, 4.0 * (double)n, st);
            b2q_launch(ctx, reduce_flat_publish_kernel<B2Q_REDUCE_UNROLL, B2Q_REDUCE_LDPOL>, (unsigned)rgrid, B2Q_THREADS, st, x,
                       sp, slot, ur, pb);
            B2Q_LAUNCH_CHECK(ctx);
            np = (int)rgrid;
        } else {
            rc = is_max ? launch_reduce_deferred<true>(ctx, slot, x, n, ur, st, &np)
                        : launch_reduce_deferred<false>(ctx, slot, x, n, ur, st, &np);
        }
        if (rc) return rc;
        B2Q_REQUIRE(np > 0, "peer path needs equally aligned float32 buffers");
        UpdateArgs u;
        memset(&u, 0, sizeof(u));
        u.mode = upd_mode;
        u.write_aux = 1; u.use_aux_as_scale = 1; u.p0 = p0; u.p1 = p1; u.aux = aux;
        const int64_t grid = b2q_flat_grid(ctx, sp.n8, B2Q_QDQ_UNROLL);
        const int rev = (ctx->reverse && n * 4 > ((long long)ctx->reverse_min_mb << 20)) ? 1 : 0;
        const bool stream_out = n * 4 > B2Q_STREAM_BYTES;
        const unsigned long long timeout_ns = (unsigned long long)(ctx->peer_timeout_ms > 0 ? ctx->peer_timeout_ms : 1) * 1000000ull;
        b2q_timed_launch tl(ctx, B2Q_KIND_QDQ_HOT, 8.0 * (double)n, st);
#define B2Q_PEER2_SWEEP(C, S) b2q_launch(ctx, qdq_peer2_kernel<C, B2Q_QDQ_UNROLL, B2Q_QDQ_LDPOL, S>, (unsigned)grid, B2Q_THREADS, st, \
            x, y, sp, pb, slot, u, qlevel, ctx->fast_div, rev, clip_with_fresh, is_max ? 1 : 0, np, (float)n, timeout_ns, published)
#define B2Q_PEER2_SWEEP_S(C) do { if (stream_out) B2Q_PEER2_SWEEP(C, 1); else B2Q_PEER2_SWEEP(C, B2Q_QDQ_STPOL); } while (0)
#define B2Q_PEER3_SWEEP(C, S, G) do {                                                                                  \
            const int64_t tiles_ = (sp.n8 + (int64_t)B2Q_THREADS * B2Q_QDQ_UNROLL - 1) / ((int64_t)B2Q_THREADS * B2Q_QDQ_UNROLL); \
            int64_t grid3_ = (tiles_ + G) / (1 + G);                                                                          \
            if (grid3_ < 1) grid3_ = 1;                                                                                        \
            b2q_prefer_shared(qdq_peer3_kernel<C, B2Q_QDQ_UNROLL, B2Q_QDQ_LDPOL, S, G>);                                       \
            b2q_launch(ctx, qdq_peer3_kernel<C, B2Q_QDQ_UNROLL, B2Q_QDQ_LDPOL, S, G>, (unsigned)grid3_, B2Q_THREADS, st, x, y,  \
                       sp, pb, slot, u, qlevel, ctx->fast_div, rev, clip_with_fresh, is_max ? 1 : 0, np, (float)n, timeout_ns, \
                       ctx->peer_stage_early);                                                                                 \
        } while (0)
#define B2Q_PEER3_SWEEP_S(C, G) do { if (stream_out) B2Q_PEER3_SWEEP(C, 1, G); else B2Q_PEER3_SWEEP(C, B2Q_QDQ_STPOL, G); } while (0)
        if (mode == 2) {
            if (clip_mode == B2Q_CLIP_SYM) B2Q_PEER3_SWEEP_S(B2Q_CLIP_SYM, 1); else B2Q_PEER3_SWEEP_S(B2Q_CLIP_NONE, 1);
        } else if (mode == 3) {
            if (clip_mode == B2Q_CLIP_SYM) B2Q_PEER3_SWEEP_S(B2Q_CLIP_SYM, 2); else B2Q_PEER3_SWEEP_S(B2Q_CLIP_NONE, 2);
        } else {
            if (clip_mode == B2Q_CLIP_SYM) B2Q_PEER2_SWEEP_S(B2Q_CLIP_SYM); else B2Q_PEER2_SWEEP_S(B2Q_CLIP_NONE);
        }
#undef B2Q_PEER3_SWEEP_S
#undef B2Q_PEER3_SWEEP
#undef B2Q_PEER2_SWEEP_S
#undef B2Q_PEER2_SWEEP
        B2Q_LAUNCH_CHECK(ctx);
        return 0;
    }
    {
        const int64_t grid = b2q_flat_grid(ctx, sp.n8, B2Q_REDUCE_UNROLL, is_max ? ctx->peer_reduce_blocks_per_sm : ctx->reduce_blocks_per_sm);
        b2q_timed_launch tl(ctx, B2Q_KIND_REDUCE_FLAT, 4.0 * (double)n, st);
        if (is_max)
            b2q_launch(ctx, reduce_peer_kernel<true, B2Q_REDUCE_UNROLL, B2Q_REDUCE_LDPOL>, (unsigned)grid, B2Q_THREADS, st,
                       x, sp, slot, aux, pb, (float)n);
        else
            b2q_launch(ctx, reduce_peer_kernel<false, B2Q_REDUCE_UNROLL, B2Q_REDUCE_LDPOL>, (unsigned)grid, B2Q_THREADS, st,
                       x, sp, slot, aux, pb, (float)n);
        B2Q_LAUNCH_CHECK(ctx);
    }
    UpdateArgs u;
    memset(&u, 0, sizeof(u));
    u.mode = upd_mode;
    u.write_aux = 1; u.use_aux_as_scale = 1; u.p0 = p0; u.p1 = p1; u.aux = aux;
    {
        const int64_t grid = b2q_flat_grid(ctx, sp.n8, B2Q_QDQ_UNROLL);
        const int rev = (ctx->reverse && n * 4 > ((long long)ctx->reverse_min_mb << 20)) ? 1 : 0;
        const bool stream_out = n * 4 > B2Q_STREAM_BYTES;   // outputs that cannot stay in L2 anyway: streaming stores
        b2q_timed_launch tl(ctx, B2Q_KIND_QDQ_HOT, 8.0 * (double)n, st);
#define B2Q_PEER_SWEEP(C, S) b2q_launch(ctx, qdq_peer_kernel<C, B2Q_QDQ_UNROLL, B2Q_QDQ_LDPOL, S>, (unsigned)grid, B2Q_THREADS, st, \
            x, y, sp, pb, (const b2q_slot*)slot, u, qlevel, ctx->fast_div, rev, clip_with_fresh)
#define B2Q_PEER_SWEEP_S(C) do { if (stream_out) B2Q_PEER_SWEEP(C, 1); else B2Q_PEER_SWEEP(C, B2Q_QDQ_STPOL); } while (0)
        if (clip_mode == B2Q_CLIP_SYM) B2Q_PEER_SWEEP_S(B2Q_CLIP_SYM); else B2Q_PEER_SWEEP_S(B2Q_CLIP_NONE);
#undef B2Q_PEER_SWEEP_S
#undef B2Q_PEER_SWEEP
        B2Q_LAUNCH_CHECK(ctx);
    }
    return 0;
}

int b2q_peer_buffer_create(b2q_ctx* ctx, int64_t bytes, void** buffer, void* ipc_handle_out) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(buffer && ipc_handle_out && bytes >= 1, "bad argument");
    void* p = nullptr;
    B2Q_CHECK_CUDA(cudaMalloc(&p, (size_t)bytes));
    B2Q_CHECK_CUDA(cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        b2q_set_error(std::string("cudaIpcGetMemHandle failed: ") + cudaGetErrorString(e));
        return 1;
    }
    memcpy(ipc_handle_out, &h, sizeof(h));
    *buffer = p;
    return 0;
}

static int peer_allreduce(b2q_ctx* ctx, bool is_max, float* const* bufs, int64_t count, float post, void* const* mailboxes,
                          int rank, int world, cudaStream_t st) {
    B2Q_REQUIRE(bufs && mailboxes && count >= 1, "bad argument");
    B2Q_REQUIRE(world >= 1 && world <= B2Q_PEER_MAX_RANKS && rank >= 0 && rank < world, "bad rank / world");
    ArPtrs p;
    memset(&p, 0, sizeof(p));
    PeerBoxes pb;
    memset(&pb, 0, sizeof(pb));
    bool aligned = true;
    for (int r = 0; r < world; ++r) {
        B2Q_REQUIRE(bufs[r] != nullptr && mailboxes[r] != nullptr, "null buffer / mailbox pointer");
        p.buf[r] = bufs[r];
        pb.box[r] = (unsigned long long*)mailboxes[r];
        aligned = aligned && ((((uintptr_t)bufs[r]) & 15) == 0);
    }
    pb.rank = rank; pb.world = world;
    pb.state = (PeerState*)((char*)mailboxes[rank] + B2Q_PEER_BOX_BYTES);
    pb.resolved = nullptr;
    // rank r owns the r-th slice, in units of four floats when every buffer is 16-byte aligned
    const int64_t unit = aligned ? 4 : 1;
    const int64_t units = (count + unit - 1) / unit;
    int64_t b = (units * rank) / world * unit, e = (units * (rank + 1)) / world * unit;
    if (e > count || rank == world - 1) e = count;
    if (b > e) b = e;
    int64_t work = (e - b + unit - 1) / unit;
    int64_t grid = (work + 255) / 256;
    const int64_t cap = ctx->peer_allreduce_blocks > 0 ? (int64_t)ctx->peer_allreduce_blocks
                                                       : (int64_t)ctx->num_sms * (ctx->peer_allreduce_blocks_per_sm > 0 ? ctx->peer_allreduce_blocks_per_sm : 2);
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    const unsigned long long timeout_ns = (unsigned long long)(ctx->peer_timeout_ms > 0 ? ctx->peer_timeout_ms : 1) * 1000000ull;
    b2q_timed_launch tl(ctx, B2Q_KIND_OTHER, 8.0 * (double)(e - b) * (double)world, st);
    if (is_max) peer_allreduce_kernel<true><<<(unsigned)grid, 256, 0, st>>>(p, pb, b, e, aligned ? 1 : 0, 1.f, timeout_ns);
    else peer_allreduce_kernel<false><<<(unsigned)grid, 256, 0, st>>>(p, pb, b, e, aligned ? 1 : 0, post, timeout_ns);
    B2Q_LAUNCH_CHECK(ctx);
    return 0;
}

int b2q_peer_allreduce_sum_f32(b2q_ctx* ctx, float* const* bufs, int64_t count, int average, void* const* mailboxes,
                               int rank, int world, void* stream) {
    B2Q_CTX(ctx);
    return peer_allreduce(ctx, false, bufs, count, average ? 1.f / (float)world : 1.f, mailboxes, rank, world,
                          (cudaStream_t)stream);
}

int b2q_peer_allreduce_max_f32(b2q_ctx* ctx, float* const* bufs, int64_t count, void* const* mailboxes, int rank, int world,
                               void* stream) {
    B2Q_CTX(ctx);
    return peer_allreduce(ctx, true, bufs, count, 1.f, mailboxes, rank, world, (cudaStream_t)stream);
}

int b2q_peer_minmax_quant_fwd_f32(b2q_ctx* ctx, int variant, const float* x, float* y, float* aux, int64_t n, int init,
                                  float ema_decay, float one_minus_decay, void* const* mailboxes, int rank, int world,
                                  uint32_t sequence, void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(variant == 0 || variant == 1, "variant must be 0 or 1");
    (void)sequence;   // kept in the ABI for logging; the authoritative counter is on the device
    // clip_grad...py:42-46 / quant_ops.py:37
    return peer_quant_fwd(ctx, true, (variant == 1 && init) ? B2Q_UPD_STORE : B2Q_UPD_EMA, ema_decay, one_minus_decay,
                          variant == 1 ? B2Q_CLIP_SYM : B2Q_CLIP_NONE, 0, 127.f, x, y, aux, n, mailboxes, rank, world,
                          (cudaStream_t)stream);
}

int b2q_peer_meanabs_quant_fwd_f32(b2q_ctx* ctx, int upd_mode, const float* x, float* y, float* aux, int64_t n, float p0,
                                   float p1, float qlevel, void* const* mailboxes, int rank, int world, void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(upd_mode == B2Q_UPD_GDRQ_ACT || upd_mode == B2Q_UPD_TWICE_STORE || upd_mode == B2Q_UPD_TWICE_EMA,
                "upd_mode must be B2Q_UPD_GDRQ_ACT, B2Q_UPD_TWICE_STORE or B2Q_UPD_TWICE_EMA");
    B2Q_REQUIRE(qlevel > 0.f, "qlevel must be positive");
    // GDRQ.py:76-86 clips with the updated alpha; fold_bn_v1_gdrq.py:65-68 clips with the batch threshold
    return peer_quant_fwd(ctx, false, upd_mode, p0, p1, B2Q_CLIP_SYM, upd_mode == B2Q_UPD_GDRQ_ACT ? 0 : 1, qlevel, x, y,
                          aux, n, mailboxes, rank, world, (cudaStream_t)stream);
}

}  // extern "C"
