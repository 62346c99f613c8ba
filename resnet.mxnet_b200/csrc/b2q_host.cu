// Host-buffer entry points: the call a framework makes when its tensors live in host memory
// (bench.py's "e2e" leg).
//
// A three-stream software pipeline keeps both PCIe directions busy:
//     h2d stream   : host x (and dy) -> device staging             continuous host-to-device traffic
//     compute strm : reduction + threshold update + QDQ sweep / backward mask on the staged tensors
//     d2h stream   : staged output -> host y / dx, aux -> host      continuous device-to-host traffic
// Staging memory is a byte-granular RING (not a fixed number of max-size slots): every call reserves exactly its
// tensor's size, so many small tensors can be in flight while the D2H of a large one drains, and the h2d stream only
// waits (stream-side, never the host) for the calls whose space it is about to overwrite.  Every call returns after
// enqueueing; b2q_host_sync() waits for all of them.  Calls that touch the same host aux array must be separated by
// b2q_host_sync().
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "b2q_common.cuh"

#define B2Q_CTX(ctx)                               \
    B2Q_REQUIRE((ctx) != nullptr, "null context"); \
    B2Q_CHECK_CUDA(cudaSetDevice((ctx)->device))

#define B2Q_HOST_AUX_SLOTS 64
#define B2Q_HOST_MIN_RING (256ll << 20)   // elements: 1 GiB per staging buffer at least

struct HostSeg {
    size_t off, len;       // elements
    cudaEvent_t done;      // recorded on the d2h stream after the call's last copy-out
};

struct HostStage {         // what one call works on
    float* a;
    float* b;
    float* c;
    float* aux;
    cudaEvent_t h2d_done, comp_done, d2h_done;
};

// Host-to-host copies (the straight-through backward of a host caller: dx <- dy, both in host memory).  Moving the
// bytes to the GPU and back would spend two PCIe transfers on an identity; a small pool of host threads copies them
// instead -- no arithmetic happens on the CPU -- while the PCIe links keep serving the forward traffic.  A job may carry
// CUDA events to wait for (device-to-host copies still landing in its source).
struct CopyJob {
    char* dst;
    const char* src;
    size_t bytes;
    cudaEvent_t wait;     // may be null
};

// memcpy with non-temporal stores: the destination is written once and not read by this thread, so the read-for-
// ownership of every destination line (a third of the memory traffic of an ordinary copy) is avoided.  512-bit stores
// where the CPU has them (one full line per store: 91 vs 81 GB/s with 16 threads on the B200 host,
// profiles/r02e_hostcopy_probe.log), 128-bit otherwise.
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
__attribute__((target("avx512f"))) static void stream_lines_512(char* dst, const char* src, size_t lines) {
    for (size_t i = 0; i + 4 <= lines; i += 4) {
        const __m512i a = _mm512_loadu_si512(src), b = _mm512_loadu_si512(src + 64);
        const __m512i c = _mm512_loadu_si512(src + 128), d = _mm512_loadu_si512(src + 192);
        _mm512_stream_si512((__m512i*)dst, a); _mm512_stream_si512((__m512i*)(dst + 64), b);
        _mm512_stream_si512((__m512i*)(dst + 128), c); _mm512_stream_si512((__m512i*)(dst + 192), d);
        src += 256; dst += 256;
    }
    for (size_t i = lines & ~(size_t)3; i < lines; ++i) {
        _mm512_stream_si512((__m512i*)dst, _mm512_loadu_si512(src));
        src += 64; dst += 64;
    }
}
static const bool kHave512 = __builtin_cpu_supports("avx512f");
#define B2Q_HAVE_STREAM_512 1
#endif

static void stream_copy(char* dst, const char* src, size_t bytes) {
#if defined(__SSE2__)
    size_t head = (64 - ((uintptr_t)dst & 63)) & 63;
    if (head > bytes) head = bytes;
    if (head) { std::memcpy(dst, src, head); dst += head; src += head; bytes -= head; }
    const size_t blocks = bytes / 64;
#ifdef B2Q_HAVE_STREAM_512
    if (kHave512) {
        stream_lines_512(dst, src, blocks);
        src += blocks * 64; dst += blocks * 64;
    } else
#endif
    if (((uintptr_t)src & 15) == 0) {
        for (size_t i = 0; i < blocks; ++i) {
            const __m128i a = _mm_load_si128((const __m128i*)(src) + 0), b = _mm_load_si128((const __m128i*)(src) + 1);
            const __m128i c = _mm_load_si128((const __m128i*)(src) + 2), d = _mm_load_si128((const __m128i*)(src) + 3);
            _mm_stream_si128((__m128i*)(dst) + 0, a); _mm_stream_si128((__m128i*)(dst) + 1, b);
            _mm_stream_si128((__m128i*)(dst) + 2, c); _mm_stream_si128((__m128i*)(dst) + 3, d);
            src += 64; dst += 64;
        }
    } else {
        for (size_t i = 0; i < blocks; ++i) {
            const __m128i a = _mm_loadu_si128((const __m128i*)(src) + 0), b = _mm_loadu_si128((const __m128i*)(src) + 1);
            const __m128i c = _mm_loadu_si128((const __m128i*)(src) + 2), d = _mm_loadu_si128((const __m128i*)(src) + 3);
            _mm_stream_si128((__m128i*)(dst) + 0, a); _mm_stream_si128((__m128i*)(dst) + 1, b);
            _mm_stream_si128((__m128i*)(dst) + 2, c); _mm_stream_si128((__m128i*)(dst) + 3, d);
            src += 64; dst += 64;
        }
    }
    _mm_sfence();
    bytes -= blocks * 64;
#endif
    if (bytes) std::memcpy(dst, src, bytes);
}

class HostCopyPool {
  public:
    explicit HostCopyPool(int device) : device_(device) {
        unsigned n = std::thread::hardware_concurrency();
        if (const char* e = std::getenv("B2Q_HOST_COPY_THREADS")) n = (unsigned)std::atoi(e);
        if (n < 1) n = 1;
        if (n > 16) n = 16;
        for (unsigned i = 0; i < n; ++i) workers_.emplace_back([this] { run(); });
    }
    ~HostCopyPool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (std::thread& t : workers_) t.join();
    }
    void submit(void* dst, const void* src, size_t bytes, cudaEvent_t wait) {
        const size_t chunk = 8u << 20;
        std::lock_guard<std::mutex> lk(m_);
        for (size_t off = 0; off < bytes; off += chunk) {
            q_.push_back({(char*)dst + off, (const char*)src + off, bytes - off < chunk ? bytes - off : chunk, wait});
            ++pending_;
        }
        cv_.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }
    int threads() const { return (int)workers_.size(); }

  private:
    void run() {
        cudaSetDevice(device_);
        for (;;) {
            CopyJob j;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [this] { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                j = q_.front();
                q_.pop_front();
            }
            if (j.wait) cudaEventSynchronize(j.wait);
            stream_copy(j.dst, j.src, j.bytes);
            {
                std::lock_guard<std::mutex> lk(m_);
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    int device_;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    std::deque<CopyJob> q_;
    size_t pending_ = 0;
    bool stop_ = false;
};

struct PendingHostWrite {   // a device-to-host copy that may still be landing in [ptr, ptr + bytes)
    const char* ptr;
    size_t bytes;
    cudaEvent_t done;
};

struct HostState {
    HostCopyPool* pool = nullptr;
    std::vector<PendingHostWrite> landing;
    float* a = nullptr;    // ring: staged inputs
    float* b = nullptr;    // ring: second input (dy) or output
    float* c = nullptr;    // ring: output of two-input ops (allocated on first use)
    size_t cap = 0, head = 0;
    std::deque<HostSeg> inflight;
    std::vector<cudaEvent_t> events;       // recycled
    float* aux = nullptr;                  // B2Q_HOST_AUX_SLOTS x B2Q_MAX_GROUPS
    cudaEvent_t aux_done[B2Q_HOST_AUX_SLOTS] = {};
    bool aux_used[B2Q_HOST_AUX_SLOTS] = {};
    unsigned calls = 0;
    cudaStream_t h2d = nullptr, comp = nullptr, d2h = nullptr;
};

static HostState* host_state(b2q_ctx* ctx) { return reinterpret_cast<HostState*>(ctx->host_state); }

static int new_event(HostState* hs, cudaEvent_t* e) {
    if (!hs->events.empty()) { *e = hs->events.back(); hs->events.pop_back(); return 0; }
    B2Q_CHECK_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    return 0;
}

static int host_init(b2q_ctx* ctx) {
    if (ctx->host_state) return 0;
    HostState* hs = new HostState();
    ctx->host_state = hs;
    B2Q_CHECK_CUDA(cudaStreamCreateWithFlags(&hs->h2d, cudaStreamNonBlocking));
    B2Q_CHECK_CUDA(cudaStreamCreateWithFlags(&hs->comp, cudaStreamNonBlocking));
    B2Q_CHECK_CUDA(cudaStreamCreateWithFlags(&hs->d2h, cudaStreamNonBlocking));
    B2Q_CHECK_CUDA(cudaMalloc(&hs->aux, sizeof(float) * B2Q_MAX_GROUPS * B2Q_HOST_AUX_SLOTS));
    for (int i = 0; i < B2Q_HOST_AUX_SLOTS; ++i)
        B2Q_CHECK_CUDA(cudaEventCreateWithFlags(&hs->aux_done[i], cudaEventDisableTiming));
    return 0;
}

static void note_landing(HostState* hs, const void* ptr, size_t bytes, cudaEvent_t done) {
    if (hs->landing.size() > 256) {   // drop the copies that have finished
        std::vector<PendingHostWrite> keep;
        for (const PendingHostWrite& w : hs->landing)
            if (cudaEventQuery(w.done) == cudaErrorNotReady) keep.push_back(w);
        (void)cudaGetLastError();
        hs->landing.swap(keep);
    }
    hs->landing.push_back({(const char*)ptr, bytes, done});
}

static int host_sync_all(HostState* hs) {
    B2Q_CHECK_CUDA(cudaStreamSynchronize(hs->h2d));
    B2Q_CHECK_CUDA(cudaStreamSynchronize(hs->comp));
    B2Q_CHECK_CUDA(cudaStreamSynchronize(hs->d2h));
    if (hs->pool) hs->pool->wait();
    hs->landing.clear();
    for (HostSeg& s : hs->inflight) hs->events.push_back(s.done);
    hs->inflight.clear();
    hs->head = 0;
    return 0;
}

// Reserve n elements of every ring buffer for one call; the h2d stream waits for the calls being overwritten.
static int take_stage(b2q_ctx* ctx, int64_t n64, bool need_c, HostStage* out) {
    int rc = host_init(ctx);
    if (rc) return rc;
    HostState* hs = host_state(ctx);
    const size_t n = ((size_t)n64 + 63) & ~(size_t)63;   // 256-byte granules keep every segment 256-bit aligned
    if (3 * n > hs->cap || (need_c && !hs->c)) {
        rc = host_sync_all(hs);
        if (rc) return rc;
        size_t cap = hs->cap;
        if (3 * n > cap) cap = 3 * n;
        if (cap < (size_t)B2Q_HOST_MIN_RING) cap = (size_t)B2Q_HOST_MIN_RING;
        if (cap != hs->cap) {
            cudaFree(hs->a); cudaFree(hs->b); cudaFree(hs->c);
            hs->a = hs->b = hs->c = nullptr;
            B2Q_CHECK_CUDA(cudaMalloc(&hs->a, sizeof(float) * cap));
            B2Q_CHECK_CUDA(cudaMalloc(&hs->b, sizeof(float) * cap));
            hs->cap = cap;
        }
        if (need_c && !hs->c) B2Q_CHECK_CUDA(cudaMalloc(&hs->c, sizeof(float) * hs->cap));
    }
    if (hs->head + n > hs->cap) {   // wrap: whatever still lives in [head, cap) must drain first
        while (!hs->inflight.empty() && hs->inflight.front().off >= hs->head) {
            B2Q_CHECK_CUDA(cudaStreamWaitEvent(hs->h2d, hs->inflight.front().done, 0));
            hs->events.push_back(hs->inflight.front().done);
            hs->inflight.pop_front();
        }
        hs->head = 0;
    }
    const size_t lo = hs->head, hi = hs->head + n;
    while (!hs->inflight.empty()) {
        const HostSeg& f = hs->inflight.front();
        if (f.off >= hi || f.off + f.len <= lo) break;   // oldest live segment does not overlap: nothing else can
        B2Q_CHECK_CUDA(cudaStreamWaitEvent(hs->h2d, f.done, 0));
        hs->events.push_back(f.done);
        hs->inflight.pop_front();
    }
    HostSeg seg;
    seg.off = lo;
    seg.len = n;
    rc = new_event(hs, &seg.done);
    if (rc) return rc;
    hs->inflight.push_back(seg);
    hs->head = hi;
    const unsigned slot = hs->calls++ % B2Q_HOST_AUX_SLOTS;
    if (hs->aux_used[slot]) B2Q_CHECK_CUDA(cudaStreamWaitEvent(hs->h2d, hs->aux_done[slot], 0));
    hs->aux_used[slot] = true;
    out->a = hs->a + lo;
    out->b = hs->b + lo;
    out->c = hs->c ? hs->c + lo : nullptr;
    out->aux = hs->aux + (size_t)slot * B2Q_MAX_GROUPS;
    out->d2h_done = seg.done;
    out->comp_done = hs->aux_done[slot];   // doubles as the aux slot's release marker (recorded last, on d2h)
    rc = new_event(hs, &out->h2d_done);
    return rc;
}

// h2d -> compute dependency, returns the event to the pool (a recorded-and-waited event may be reused at once)
static int chain(HostState* hs, cudaEvent_t e, cudaStream_t from, cudaStream_t to) {
    B2Q_CHECK_CUDA(cudaEventRecord(e, from));
    B2Q_CHECK_CUDA(cudaStreamWaitEvent(to, e, 0));
    return 0;
}

static int finish(HostState* hs, HostStage& s) {
    B2Q_CHECK_CUDA(cudaEventRecord(s.d2h_done, hs->d2h));
    B2Q_CHECK_CUDA(cudaEventRecord(s.comp_done, hs->d2h));
    hs->events.push_back(s.h2d_done);
    return 0;
}

int b2q_host_release(b2q_ctx* ctx) {
    if (!ctx || !ctx->host_state) return 0;
    HostState* hs = host_state(ctx);
    if (hs->h2d) { cudaStreamSynchronize(hs->h2d); cudaStreamDestroy(hs->h2d); }
    if (hs->comp) { cudaStreamSynchronize(hs->comp); cudaStreamDestroy(hs->comp); }
    if (hs->d2h) { cudaStreamSynchronize(hs->d2h); cudaStreamDestroy(hs->d2h); }
    if (hs->pool) { hs->pool->wait(); delete hs->pool; hs->pool = nullptr; }
    cudaFree(hs->a); cudaFree(hs->b); cudaFree(hs->c); cudaFree(hs->aux);
    for (HostSeg& s : hs->inflight) cudaEventDestroy(s.done);
    for (cudaEvent_t e : hs->events) cudaEventDestroy(e);
    for (cudaEvent_t e : hs->aux_done) if (e) cudaEventDestroy(e);
    delete hs;
    ctx->host_state = nullptr;
    return 0;
}

extern "C" {

int b2q_host_sync(b2q_ctx* ctx) {
    B2Q_CTX(ctx);
    if (!ctx->host_state) return 0;
    return host_sync_all(host_state(ctx));
}

int b2q_minmax_quant_fwd_host_f32(b2q_ctx* ctx, int variant, const float* host_x, float* host_y, float* host_aux,
                                  int64_t rows, int64_t cols, int is_weight, int per_channel, int is_train, int init,
                                  float ema_decay, float one_minus_decay) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(host_x && host_y && host_aux && rows >= 1 && cols >= 1, "bad argument");
    const int64_t n = rows * cols;
    const int64_t naux = (per_channel && is_weight) ? rows : 1;
    B2Q_REQUIRE(naux <= B2Q_MAX_GROUPS, "too many channels");
    HostStage s;
    int rc = take_stage(ctx, n, false, &s);
    if (rc) return rc;
    HostState* hs = host_state(ctx);
    B2Q_CHECK_CUDA(cudaMemcpyAsync(s.aux, host_aux, sizeof(float) * naux, cudaMemcpyHostToDevice, hs->h2d));
    B2Q_CHECK_CUDA(cudaMemcpyAsync(s.a, host_x, sizeof(float) * n, cudaMemcpyHostToDevice, hs->h2d));
    if ((rc = chain(hs, s.h2d_done, hs->h2d, hs->comp))) return rc;
    rc = b2q_minmax_quant_fwd_f32(ctx, variant, s.a, s.b, s.aux, rows, cols, is_weight, per_channel, is_train, init,
                                  ema_decay, one_minus_decay, B2Q_REQ_WRITE, hs->comp);
    if (rc) return rc;
    cudaEvent_t e;
    if ((rc = new_event(hs, &e))) return rc;
    if ((rc = chain(hs, e, hs->comp, hs->d2h))) return rc;
    hs->events.push_back(e);
    B2Q_CHECK_CUDA(cudaMemcpyAsync(host_y, s.b, sizeof(float) * n, cudaMemcpyDeviceToHost, hs->d2h));
    B2Q_CHECK_CUDA(cudaMemcpyAsync(host_aux, s.aux, sizeof(float) * naux, cudaMemcpyDeviceToHost, hs->d2h));
    rc = finish(hs, s);
    note_landing(hs, host_y, sizeof(float) * n, s.d2h_done);
    return rc;
}

int b2q_ste_bwd_host_f32(b2q_ctx* ctx, const float* host_dy, float* host_dx, int64_t n) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(host_dy && host_dx && n >= 1, "bad argument");
    if (host_dy == host_dx) return 0;
    if (ctx->host_ste_copy) {
        // dx <- dy between two host buffers: an identity has no business crossing PCIe twice.  Copied by the host copy
        // pool (asynchronous like every host-buffer call; b2q_host_sync() waits for it), after any device-to-host copy
        // that is still landing in dy (e.g. dy is the output of an earlier forward call).
        int rc = host_init(ctx);
        if (rc) return rc;
        HostState* hs = host_state(ctx);
        if (!hs->pool) hs->pool = new HostCopyPool(ctx->device);
        const char* lo = (const char*)host_dy;
        const char* hi = lo + sizeof(float) * (size_t)n;
        cudaEvent_t wait = nullptr;
        for (size_t i = hs->landing.size(); i-- > 0;) {   // newest first: events on the d2h stream complete in order
            const PendingHostWrite& w = hs->landing[i];
            if (w.ptr < hi && lo < w.ptr + w.bytes) { wait = w.done; break; }
        }
        hs->pool->submit(host_dx, host_dy, sizeof(float) * (size_t)n, wait);
        return 0;
    }
    HostStage s;
    int rc = take_stage(ctx, n, false, &s);
    if (rc) return rc;
    HostState* hs = host_state(ctx);
    B2Q_CHECK_CUDA(cudaMemcpyAsync(s.a, host_dy, sizeof(float) * n, cudaMemcpyHostToDevice, hs->h2d));
    if ((rc = chain(hs, s.h2d_done, hs->h2d, hs->comp))) return rc;
    rc = b2q_ste_bwd_f32(ctx, s.a, s.b, n, B2Q_REQ_WRITE, hs->comp);
    if (rc) return rc;
    cudaEvent_t e;
    if ((rc = new_event(hs, &e))) return rc;
    if ((rc = chain(hs, e, hs->comp, hs->d2h))) return rc;
    hs->events.push_back(e);
    B2Q_CHECK_CUDA(cudaMemcpyAsync(host_dx, s.b, sizeof(float) * n, cudaMemcpyDeviceToHost, hs->d2h));
    return finish(hs, s);
}

int b2q_clipgrad_bwd_host_f32(b2q_ctx* ctx, const float* host_x, const float* host_dy, float* host_dx,
                              const float* host_aux, int64_t n) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(host_x && host_dy && host_dx && host_aux && n >= 1, "bad argument");
    HostStage s;
    int rc = take_stage(ctx, n, true, &s);
    if (rc) return rc;
    HostState* hs = host_state(ctx);
    B2Q_CHECK_CUDA(cudaMemcpyAsync(s.aux, host_aux, sizeof(float), cudaMemcpyHostToDevice, hs->h2d));
    B2Q_CHECK_CUDA(cudaMemcpyAsync(s.a, host_x, sizeof(float) * n, cudaMemcpyHostToDevice, hs->h2d));
    B2Q_CHECK_CUDA(cudaMemcpyAsync(s.b, host_dy, sizeof(float) * n, cudaMemcpyHostToDevice, hs->h2d));
    if ((rc = chain(hs, s.h2d_done, hs->h2d, hs->comp))) return rc;
    rc = b2q_clipgrad_bwd_f32(ctx, s.a, s.b, s.c, s.aux, n, hs->comp);
    if (rc) return rc;
    cudaEvent_t e;
    if ((rc = new_event(hs, &e))) return rc;
    if ((rc = chain(hs, e, hs->comp, hs->d2h))) return rc;
    hs->events.push_back(e);
    B2Q_CHECK_CUDA(cudaMemcpyAsync(host_dx, s.c, sizeof(float) * n, cudaMemcpyDeviceToHost, hs->d2h));
    return finish(hs, s);
}

}  // extern "C"
