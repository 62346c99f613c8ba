// What can the host side of the e2e leg deliver?  Host-to-host copy bandwidth of pinned memory against the number of
// copying threads and the store flavour (SSE2 / AVX2 / AVX-512 non-temporal, plain memcpy), alone, next to a GPU copy
// engine doing the same job over PCIe (cudaMemcpyAsync host->host), and next to saturated H2D + D2H traffic.
//   nvcc -O2 -std=c++17 -Xcompiler -march=x86-64-v2 -o tools/hostcopy_probe tools/hostcopy_probe.cu
#include <cuda_runtime.h>
#include <immintrin.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

static double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static void copy_sse2(char* d, const char* s, size_t n) {
    for (size_t i = 0; i < n; i += 64) {
        __m128i a = _mm_load_si128((const __m128i*)(s + i)), b = _mm_load_si128((const __m128i*)(s + i + 16));
        __m128i c = _mm_load_si128((const __m128i*)(s + i + 32)), e = _mm_load_si128((const __m128i*)(s + i + 48));
        _mm_stream_si128((__m128i*)(d + i), a); _mm_stream_si128((__m128i*)(d + i + 16), b);
        _mm_stream_si128((__m128i*)(d + i + 32), c); _mm_stream_si128((__m128i*)(d + i + 48), e);
    }
    _mm_sfence();
}

__attribute__((target("avx2"))) static void copy_avx2(char* d, const char* s, size_t n) {
    for (size_t i = 0; i < n; i += 128) {
        __m256i a = _mm256_load_si256((const __m256i*)(s + i)), b = _mm256_load_si256((const __m256i*)(s + i + 32));
        __m256i c = _mm256_load_si256((const __m256i*)(s + i + 64)), e = _mm256_load_si256((const __m256i*)(s + i + 96));
        _mm256_stream_si256((__m256i*)(d + i), a); _mm256_stream_si256((__m256i*)(d + i + 32), b);
        _mm256_stream_si256((__m256i*)(d + i + 64), c); _mm256_stream_si256((__m256i*)(d + i + 96), e);
    }
    _mm_sfence();
}

__attribute__((target("avx2"))) static void copy_avx2_pf(char* d, const char* s, size_t n) {
    for (size_t i = 0; i < n; i += 128) {
        _mm_prefetch(s + i + 2048, _MM_HINT_NTA);
        _mm_prefetch(s + i + 2048 + 64, _MM_HINT_NTA);
        __m256i a = _mm256_load_si256((const __m256i*)(s + i)), b = _mm256_load_si256((const __m256i*)(s + i + 32));
        __m256i c = _mm256_load_si256((const __m256i*)(s + i + 64)), e = _mm256_load_si256((const __m256i*)(s + i + 96));
        _mm256_stream_si256((__m256i*)(d + i), a); _mm256_stream_si256((__m256i*)(d + i + 32), b);
        _mm256_stream_si256((__m256i*)(d + i + 64), c); _mm256_stream_si256((__m256i*)(d + i + 96), e);
    }
    _mm_sfence();
}

__attribute__((target("avx512f"))) static void copy_avx512(char* d, const char* s, size_t n) {
    for (size_t i = 0; i < n; i += 256) {
        __m512i a = _mm512_load_si512(s + i), b = _mm512_load_si512(s + i + 64);
        __m512i c = _mm512_load_si512(s + i + 128), e = _mm512_load_si512(s + i + 192);
        _mm512_stream_si512((__m512i*)(d + i), a); _mm512_stream_si512((__m512i*)(d + i + 64), b);
        _mm512_stream_si512((__m512i*)(d + i + 128), c); _mm512_stream_si512((__m512i*)(d + i + 192), e);
    }
    _mm_sfence();
}

static void copy_memcpy(char* d, const char* s, size_t n) { std::memcpy(d, s, n); }

typedef void (*copy_fn)(char*, const char*, size_t);

// nthreads workers pull 8 MB chunks of one big copy from an atomic counter; returns seconds
static double threaded_copy(copy_fn fn, char* dst, const char* src, size_t bytes, int nthreads, size_t chunk = 8u << 20) {
    std::atomic<size_t> next{0};
    const size_t nchunks = (bytes + chunk - 1) / chunk;
    const double t0 = now();
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([&] {
            for (;;) {
                const size_t c = next.fetch_add(1);
                if (c >= nchunks) return;
                const size_t off = c * chunk, len = bytes - off < chunk ? bytes - off : chunk;
                fn(dst + off, src + off, len);
            }
        });
    for (auto& t : th) t.join();
    return now() - t0;
}

int main(int argc, char** argv) {
    const size_t bytes = (size_t)(argc > 1 ? atoi(argv[1]) : 2048) << 20;
    const int hw = (int)std::thread::hardware_concurrency();
    printf("hardware_concurrency %d, buffer %zu MB\n", hw, bytes >> 20);
    char *src, *dst, *src2, *dst2;
    cudaSetDevice(0);
    if (cudaHostAlloc(&src, bytes, cudaHostAllocDefault) || cudaHostAlloc(&dst, bytes, cudaHostAllocDefault) ||
        cudaHostAlloc(&src2, bytes, cudaHostAllocDefault) || cudaHostAlloc(&dst2, bytes, cudaHostAllocDefault)) {
        printf("cudaHostAlloc failed\n");
        return 1;
    }
    std::memset(src, 1, bytes); std::memset(dst, 2, bytes); std::memset(src2, 3, bytes); std::memset(dst2, 4, bytes);
    char *da, *db;
    cudaMalloc(&da, bytes); cudaMalloc(&db, bytes);
    cudaStream_t s1, s2, s3;
    cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&s3, cudaStreamNonBlocking);
    const bool has_avx2 = __builtin_cpu_supports("avx2"), has_512 = __builtin_cpu_supports("avx512f");
    struct { const char* name; copy_fn fn; bool ok; } kinds[] = {
        {"sse2-nt", copy_sse2, true}, {"avx2-nt", copy_avx2, has_avx2}, {"avx2-nt+prefetch", copy_avx2_pf, has_avx2},
        {"avx512-nt", copy_avx512, has_512}, {"memcpy", copy_memcpy, true}};
    const int counts[] = {1, 2, 4, 8, 12, 16, 20, 24, 32, 48, 64};
    printf("== host-to-host copy alone: GB/s copied (memory traffic is twice that) ==\n");
    for (auto& k : kinds) {
        if (!k.ok) continue;
        printf("%-18s", k.name);
        for (int n : counts) {
            if (n > 2 * hw) break;
            threaded_copy(k.fn, dst, src, bytes / 4, n);
            double best = 1e9;
            for (int r = 0; r < 3; ++r) { double t = threaded_copy(k.fn, dst, src, bytes, n); if (t < best) best = t; }
            printf(" %d:%.1f", n, bytes / best / 1e9);
            fflush(stdout);
        }
        printf("\n");
    }
    // GPU copy engine, host -> host
    {
        cudaMemcpyAsync(dst2, src2, bytes, cudaMemcpyHostToHost, s3);
        cudaStreamSynchronize(s3);
        double t0 = now();
        cudaMemcpyAsync(dst2, src2, bytes, cudaMemcpyHostToHost, s3);
        cudaStreamSynchronize(s3);
        double t = now() - t0;
        printf("== cudaMemcpyAsync host->host (pinned): %.1f GB/s ==\n", bytes / t / 1e9);
        t0 = now();
        cudaMemcpyAsync(da, src2, bytes, cudaMemcpyHostToDevice, s1);
        cudaMemcpyAsync(dst2, db, bytes, cudaMemcpyDeviceToHost, s2);
        cudaStreamSynchronize(s1); cudaStreamSynchronize(s2);
        t = now() - t0;
        printf("== H2D + D2H concurrently: %.1f GB/s per direction ==\n", bytes / t / 1e9);
    }
    copy_fn best_fn = has_512 ? copy_avx512 : has_avx2 ? copy_avx2 : copy_sse2;
    printf("== host copy (best vector flavour) while H2D + D2H run: copy GB/s | pcie GB/s per direction ==\n");
    for (int n : counts) {
        if (n > 2 * hw) break;
        std::atomic<bool> stop{false};
        std::atomic<long> rounds{0};
        std::thread pcie([&] {
            while (!stop.load()) {
                cudaMemcpyAsync(da, src2, bytes, cudaMemcpyHostToDevice, s1);
                cudaMemcpyAsync(dst2, db, bytes, cudaMemcpyDeviceToHost, s2);
                cudaStreamSynchronize(s1); cudaStreamSynchronize(s2);
                rounds.fetch_add(1);
            }
        });
        while (rounds.load() < 1) std::this_thread::yield();
        const long r0 = rounds.load();
        const double t0 = now();
        double tc = 0;
        int reps = 0;
        while (rounds.load() < r0 + 3) { tc += threaded_copy(best_fn, dst, src, bytes, n); ++reps; }
        const double t = now() - t0;
        const long r1 = rounds.load();
        stop.store(true);
        pcie.join();
        printf(" threads %2d: copy %.1f | pcie %.1f\n", n, reps * (double)bytes / tc / 1e9, (r1 - r0) * (double)bytes / t / 1e9);
        fflush(stdout);
    }
    printf("== host copy threads + GPU copy engine (host->host) sharing one job ==\n");
    for (int n : {8, 16, 24, 32}) {
        if (n > 2 * hw) break;
        for (double frac : {0.0, 0.2, 0.3, 0.4}) {
            const size_t gpu_bytes = ((size_t)(bytes * frac)) & ~(size_t)4095;
            const double t0 = now();
            if (gpu_bytes) cudaMemcpyAsync(dst, src, gpu_bytes, cudaMemcpyHostToHost, s3);
            threaded_copy(best_fn, dst + gpu_bytes, src + gpu_bytes, bytes - gpu_bytes, n);
            cudaStreamSynchronize(s3);
            const double t = now() - t0;
            printf(" threads %2d gpu share %.1f: %.1f GB/s\n", n, frac, bytes / t / 1e9);
        }
    }
    return 0;
}
