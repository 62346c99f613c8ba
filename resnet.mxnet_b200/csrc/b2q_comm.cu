// One process driving N devices (the reference's layout: train.py:34, core/solver.py:58-61 -- a Module executor group and
// a KVStore, no launcher): the two exchanges of data-parallel training over peer memory, without torch.distributed,
// NCCL or CUDA IPC.
//
//   b2q_comm_create          enables peer access between the contexts' devices and gives every rank a mailbox; the
//                            per-rank mailbox tables feed the fused threshold kernels (b2q_peer_*_quant_fwd_f32), which
//                            need nothing else -- each device's sweep reads the statistics its peers stored into its
//                            mailbox over NVLink
//   b2q_comm_allreduce_*     element-wise max / sum of one float32 buffer per rank, in place, result bit-identical on
//                            every rank.  Rank r owns the r-th slice: its kernel reads that slice from every rank
//                            (peer loads over NVLink), combines the values in rank order (so every rank holds the same
//                            bits) and stores the result into every rank's buffer (peer stores) -- reduce-scatter and
//                            all-gather in one kernel, (n-1)/n of the buffer in each direction per device like a ring,
//                            but one hop.  Ordering across devices is by CUDA events (record on every stream, every
//                            stream waits for all of them) before and after, so the calls are asynchronous and
//                            stream-ordered like everything else in the library.
// KVStore semantics kept: gradients are summed over devices (kvstore 'device', solver.py:121); `average` divides by n.
#include <cstring>
#include <vector>

#include "b2q_common.cuh"

#define B2Q_COMM_MAX_RANKS 16

struct b2q_comm {
    int n = 0;
    b2q_ctx* ctx[B2Q_COMM_MAX_RANKS] = {};
    void* mailbox[B2Q_COMM_MAX_RANKS] = {};
    void* table[B2Q_COMM_MAX_RANKS][B2Q_COMM_MAX_RANKS] = {};   // table[r] = what rank r passes as `mailboxes`
    cudaEvent_t ev_in[B2Q_COMM_MAX_RANKS] = {}, ev_out[B2Q_COMM_MAX_RANKS] = {};
};

struct CommPtrs {
    float* buf[B2Q_COMM_MAX_RANKS];
};

// slice owner: reads its slice from every rank, combines in rank order, writes it to every rank
template <bool IS_MAX>
__global__ void __launch_bounds__(256)
comm_allreduce_kernel(CommPtrs p, int n_ranks, int64_t begin, int64_t end, float post_scale) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const bool vec = ((begin | end) & 3) == 0;
    if (vec) {
        for (int64_t i = (begin >> 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (end >> 2); i += stride) {
            float4 acc = reinterpret_cast<const float4*>(p.buf[0])[i];
            for (int r = 1; r < n_ranks; ++r) {
                const float4 v = reinterpret_cast<const float4*>(p.buf[r])[i];
                if (IS_MAX) {
                    acc.x = fmax_nan(acc.x, v.x); acc.y = fmax_nan(acc.y, v.y); acc.z = fmax_nan(acc.z, v.z); acc.w = fmax_nan(acc.w, v.w);
                } else {
                    acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y); acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
                }
            }
            if (!IS_MAX && post_scale != 1.f) {
                acc.x = __fmul_rn(acc.x, post_scale); acc.y = __fmul_rn(acc.y, post_scale);
                acc.z = __fmul_rn(acc.z, post_scale); acc.w = __fmul_rn(acc.w, post_scale);
            }
            for (int r = 0; r < n_ranks; ++r) reinterpret_cast<float4*>(p.buf[r])[i] = acc;
        }
        return;
    }
    for (int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += stride) {
        float acc = p.buf[0][i];
        for (int r = 1; r < n_ranks; ++r) acc = IS_MAX ? fmax_nan(acc, p.buf[r][i]) : __fadd_rn(acc, p.buf[r][i]);
        if (!IS_MAX && post_scale != 1.f) acc = __fmul_rn(acc, post_scale);
        for (int r = 0; r < n_ranks; ++r) p.buf[r][i] = acc;
    }
}

extern "C" int b2q_peer_mailbox_bytes(void);

static int comm_allreduce(b2q_comm* c, bool is_max, float* const* bufs, int64_t count, void* const* streams, float post) {
    B2Q_REQUIRE(c && bufs && streams && count >= 1, "bad argument");
    const int n = c->n;
    CommPtrs p;
    memset(&p, 0, sizeof(p));
    bool aligned = true;
    for (int r = 0; r < n; ++r) {
        B2Q_REQUIRE(bufs[r] != nullptr, "null buffer");
        p.buf[r] = bufs[r];
        aligned = aligned && ((((uintptr_t)bufs[r]) & 15) == 0);
    }
    // inputs complete on every device before anyone reads them
    for (int r = 0; r < n; ++r) {
        B2Q_CHECK_CUDA(cudaSetDevice(c->ctx[r]->device));
        B2Q_CHECK_CUDA(cudaEventRecord(c->ev_in[r], (cudaStream_t)streams[r]));
    }
    for (int r = 0; r < n; ++r) {
        B2Q_CHECK_CUDA(cudaSetDevice(c->ctx[r]->device));
        for (int q = 0; q < n; ++q)
            if (q != r) B2Q_CHECK_CUDA(cudaStreamWaitEvent((cudaStream_t)streams[r], c->ev_in[q], 0));
    }
    // slices in units of 4 floats when every buffer is 16-byte aligned
    const int64_t unit = aligned ? 4 : 1;
    const int64_t units = (count + unit - 1) / unit;
    for (int r = 0; r < n; ++r) {
        int64_t b = (units * r) / n * unit, e = (units * (r + 1)) / n * unit;
        if (e > count) e = count;
        if (r == n - 1) e = count;
        if (b >= e) continue;
        const int64_t len = e - b;
        // the tail slice may end off a multiple of four: split it so that the vector body stays aligned
        const int64_t body_end = aligned ? b + (len & ~(int64_t)3) : e;
        B2Q_CHECK_CUDA(cudaSetDevice(c->ctx[r]->device));
        cudaStream_t st = (cudaStream_t)streams[r];
        for (int part = 0; part < 2; ++part) {
            const int64_t pb = part == 0 ? b : body_end, pe = part == 0 ? body_end : e;
            if (pb >= pe) continue;
            int64_t work = (pe - pb) / (part == 0 ? unit : 1);
            int64_t grid = (work + 255) / 256;
            const int64_t cap = (int64_t)c->ctx[r]->num_sms * 8;
            if (grid > cap) grid = cap;
            if (grid < 1) grid = 1;
            if (is_max) comm_allreduce_kernel<true><<<(unsigned)grid, 256, 0, st>>>(p, n, pb, pe, 1.f);
            else comm_allreduce_kernel<false><<<(unsigned)grid, 256, 0, st>>>(p, n, pb, pe, post);
            B2Q_LAUNCH_CHECK(c->ctx[r]);
        }
    }
    // every slice written everywhere before any rank continues
    for (int r = 0; r < n; ++r) {
        B2Q_CHECK_CUDA(cudaSetDevice(c->ctx[r]->device));
        B2Q_CHECK_CUDA(cudaEventRecord(c->ev_out[r], (cudaStream_t)streams[r]));
    }
    for (int r = 0; r < n; ++r) {
        B2Q_CHECK_CUDA(cudaSetDevice(c->ctx[r]->device));
        for (int q = 0; q < n; ++q)
            if (q != r) B2Q_CHECK_CUDA(cudaStreamWaitEvent((cudaStream_t)streams[r], c->ev_out[q], 0));
    }
    return 0;
}

extern "C" {

int b2q_comm_create(b2q_ctx* const* ctxs, int n, b2q_comm** out) {
    B2Q_REQUIRE(ctxs && out && n >= 1 && n <= B2Q_COMM_MAX_RANKS, "bad argument (1..16 contexts)");
    for (int r = 0; r < n; ++r) {
        B2Q_REQUIRE(ctxs[r] != nullptr, "null context");
        for (int q = 0; q < r; ++q) B2Q_REQUIRE(ctxs[q]->device != ctxs[r]->device, "one context per device");
    }
    b2q_comm* c = new b2q_comm();
    c->n = n;
    for (int r = 0; r < n; ++r) c->ctx[r] = ctxs[r];
    for (int r = 0; r < n; ++r) {
        cudaError_t e = cudaSetDevice(ctxs[r]->device);
        for (int q = 0; q < n && e == cudaSuccess; ++q) {
            if (q == r) continue;
            int can = 0;
            e = cudaDeviceCanAccessPeer(&can, ctxs[r]->device, ctxs[q]->device);
            if (e == cudaSuccess && !can) {
                delete c;
                b2q_set_error("b2q_comm_create: devices cannot access each other's memory (no NVLink / P2P)");
                return 3;
            }
            if (e == cudaSuccess) {
                e = cudaDeviceEnablePeerAccess(ctxs[q]->device, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); e = cudaSuccess; }
            }
        }
        if (e == cudaSuccess) e = cudaMalloc(&c->mailbox[r], (size_t)b2q_peer_mailbox_bytes());
        if (e == cudaSuccess) e = cudaMemset(c->mailbox[r], 0, (size_t)b2q_peer_mailbox_bytes());
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_in[r], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_out[r], cudaEventDisableTiming);
        if (e != cudaSuccess) {
            b2q_set_error(std::string("b2q_comm_create failed: ") + cudaGetErrorString(e));
            delete c;   // (device memory of a half-built communicator is reclaimed with the process)
            return 1;
        }
    }
    for (int r = 0; r < n; ++r)
        for (int q = 0; q < n; ++q) c->table[r][q] = c->mailbox[q];   // one address space: every rank sees every mailbox
    *out = c;
    return 0;
}

int b2q_comm_destroy(b2q_comm* c) {
    if (!c) return 0;
    for (int r = 0; r < c->n; ++r) {
        cudaSetDevice(c->ctx[r]->device);
        cudaDeviceSynchronize();
        if (c->mailbox[r]) cudaFree(c->mailbox[r]);
        if (c->ev_in[r]) cudaEventDestroy(c->ev_in[r]);
        if (c->ev_out[r]) cudaEventDestroy(c->ev_out[r]);
    }
    delete c;
    return 0;
}

int b2q_comm_size(b2q_comm* c) { return c ? c->n : -1; }

int b2q_comm_mailboxes(b2q_comm* c, int rank, void* const** mailboxes) {
    B2Q_REQUIRE(c && mailboxes && rank >= 0 && rank < c->n, "bad argument");
    *mailboxes = c->table[rank];
    return 0;
}

int b2q_comm_allreduce_max_f32(b2q_comm* c, float* const* bufs, int64_t count, void* const* streams) {
    return comm_allreduce(c, true, bufs, count, streams, 1.f);
}

int b2q_comm_allreduce_sum_f32(b2q_comm* c, float* const* bufs, int64_t count, int average, void* const* streams) {
    B2Q_REQUIRE(c != nullptr, "null communicator");
    return comm_allreduce(c, false, bufs, count, streams, average ? 1.f / (float)c->n : 1.f);
}

}  // extern "C"
