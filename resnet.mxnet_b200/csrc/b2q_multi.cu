// Multi-tensor launches: every weight tensor of a network quantised in TWO launches (max|w| for all tensors, then
// the QDQ sweep for all tensors) and their straight-through backward in ONE, instead of three launches per tensor.
// The 54 / 28 / 105 weight tensors of the reference networks hold ~1% of the bytes of a step but, launched one by
// one, ~7% of its time (profiles/r01a_launches_summary.md): they are launch-latency-bound.
//
// A plan is built once (weight pointers are stable across steps): a device table of tensors and a device table of
// work items.  Item kinds: RANGE = a slice of <= 8192 elements of one group (whole tensor, or one long row);
// ROWS = up to 64 consecutive short rows of a per-channel tensor, one warp per row (depthwise 3x3: 9 elements).
// max is combined with atomicMax on (epoch << 32 | bit pattern of the non-negative float): exact, order-independent and
// never in need of a reset, because the epoch lives on the device and is advanced by the sweep kernel that consumes
// the statistics -- a CUDA graph replaying the two launches keeps working on fresh tags.
#include <cstring>
#include <vector>

#include "b2q_common.cuh"
#include "b2q_qdq.cuh"

#define B2Q_CTX(ctx)                               \
    B2Q_REQUIRE((ctx) != nullptr, "null context"); \
    B2Q_CHECK_CUDA(cudaSetDevice((ctx)->device))

#define MT_ITEM_ELEMS 8192
#define MT_ROWS_PER_ITEM 64
#define MT_SHORT_ROW 512

struct MtTensor {
    const float* x;
    float* y;
    float* aux;
    const float* dy;
    float* dx;
    long long rows, cols;
    int per_channel;
    int gbase;       // index of this tensor's first group in the statistic buffers
};

struct MtItem {
    int tensor;
    int kind;        // 0 RANGE, 1 ROWS
    int group;       // RANGE: group (row) inside the tensor
    int first;       // RANGE: 1 if this item starts its group (it owns the aux write)
    int pfirst;      // RANGE: index of the first item of this group ...
    int pcount;      // ... and how many items the group has (their double partials are combined in order)
    long long a, b;  // RANGE: element range inside the group; ROWS: row range
};

struct b2q_multi_plan {
    int device;
    int n_tensors, n_items, n_groups;
    MtTensor* d_tensors;
    MtItem* d_items;
    unsigned long long* d_stat;
    unsigned int* d_epoch;
    double* d_partial;   // one per item (mean-based operators)
    long long elements;
    int has_grad;
};

__device__ __forceinline__ const float* mt_group_base(const MtTensor& t, int group) {
    return t.x + (t.per_channel ? (long long)group * t.cols : 0);
}

__global__ void __launch_bounds__(B2Q_THREADS)
mt_absmax_kernel(const MtTensor* __restrict__ tensors, const MtItem* __restrict__ items,
                 unsigned long long* __restrict__ stat, const unsigned int* __restrict__ epoch) {
    b2q_pdl_sync();
    __shared__ double smem[32];
    const unsigned long long tag = (unsigned long long)(*epoch + 1u) << 32;
    const MtItem it = items[blockIdx.x];
    const MtTensor t = tensors[it.tensor];
    if (it.kind == 0) {
        const float* base = mt_group_base(t, it.group);
        float m = 0.f;
        const bool vec = ((((uintptr_t)(base + it.a)) & 15) == 0) && (((it.b - it.a) & 3) == 0);
        if (vec) {
            const float4* p = reinterpret_cast<const float4*>(base + it.a);
            const long long n4 = (it.b - it.a) >> 2;
            for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
                const float4 v = p[i];
                m = fmax_nan(fmax_nan(m, fabsf(v.x)), fmax_nan(fabsf(v.y), fmax_nan(fabsf(v.z), fabsf(v.w))));
            }
        } else {
            for (long long i = it.a + threadIdx.x; i < it.b; i += blockDim.x) m = fmax_nan(m, fabsf(base[i]));
        }
        const float r = (float)block_reduce<true>((double)m, smem);
        if (threadIdx.x == 0) atomicMax(stat + t.gbase + it.group, tag | __float_as_uint(r));
    } else {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
        for (long long row = it.a + wid; row < it.b; row += nw) {
            const float* p = t.x + row * t.cols;
            float m = 0.f;
            for (long long c = lane; c < t.cols; c += 32) m = fmax_nan(m, fabsf(p[c]));
            m = warp_max(m);
            if (lane == 0) stat[t.gbase + row] = tag | __float_as_uint(m);
        }
    }
}

__global__ void __launch_bounds__(B2Q_THREADS)
mt_qdq_kernel(const MtTensor* __restrict__ tensors, const MtItem* __restrict__ items,
              const unsigned long long* __restrict__ stat, unsigned int* __restrict__ epoch, int from_stat, int write_aux,
              int fast) {
    b2q_pdl_sync();
    if (from_stat && blockIdx.x == 0 && threadIdx.x == 0) *epoch = (unsigned int)(stat[0] >> 32);   // consume the tag
    const MtItem it = items[blockIdx.x];
    const MtTensor t = tensors[it.tensor];
    if (it.kind == 0) {
        const float T = from_stat ? __uint_as_float((unsigned int)(stat[t.gbase + it.group] & 0xffffffffull)) : t.aux[it.group];
        const QScale s = make_qscale(T, 127.f, fast != 0);
        if (write_aux && it.first && threadIdx.x == 0) t.aux[it.group] = T;
        const long long off = (t.per_channel ? (long long)it.group * t.cols : 0);
        const float* xb = t.x + off;
        float* yb = t.y + off;
        const bool vec = ((((uintptr_t)(xb + it.a)) & 15) == 0) && ((((uintptr_t)(yb + it.a)) & 15) == 0) &&
                         (((it.b - it.a) & 3) == 0);
        if (vec) {
            const float4* p = reinterpret_cast<const float4*>(xb + it.a);
            float4* o = reinterpret_cast<float4*>(yb + it.a);
            const long long n4 = (it.b - it.a) >> 2;
            for (long long i = threadIdx.x; i < n4; i += blockDim.x) {
                const float4 v = p[i];
                float4 r;
                r.x = __fmul_rn(quant_code(v.x, s), s.q); r.y = __fmul_rn(quant_code(v.y, s), s.q);
                r.z = __fmul_rn(quant_code(v.z, s), s.q); r.w = __fmul_rn(quant_code(v.w, s), s.q);
                o[i] = r;
            }
        } else {
            for (long long i = it.a + threadIdx.x; i < it.b; i += blockDim.x) yb[i] = __fmul_rn(quant_code(xb[i], s), s.q);
        }
    } else {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
        for (long long row = it.a + wid; row < it.b; row += nw) {
            const float T = from_stat ? __uint_as_float((unsigned int)(stat[t.gbase + row] & 0xffffffffull)) : t.aux[row];
            const QScale s = make_qscale(T, 127.f, fast != 0);
            if (write_aux && lane == 0) t.aux[row] = T;
            const float* p = t.x + row * t.cols;
            float* o = t.y + row * t.cols;
            for (long long c = lane; c < t.cols; c += 32) o[c] = __fmul_rn(quant_code(p[c], s), s.q);
        }
    }
}

__global__ void __launch_bounds__(B2Q_THREADS)
mt_copy_kernel(const MtTensor* __restrict__ tensors, const MtItem* __restrict__ items) {
    b2q_pdl_sync();
    const MtItem it = items[blockIdx.x];
    const MtTensor t = tensors[it.tensor];
    long long a, b;
    if (it.kind == 0) {
        const long long off = (t.per_channel ? (long long)it.group * t.cols : 0);
        a = off + it.a; b = off + it.b;
    } else {
        a = it.a * t.cols; b = it.b * t.cols;
    }
    const bool vec = ((((uintptr_t)(t.dy + a)) & 15) == 0) && ((((uintptr_t)(t.dx + a)) & 15) == 0) && (((b - a) & 3) == 0);
    if (vec) {
        const float4* p = reinterpret_cast<const float4*>(t.dy + a);
        float4* o = reinterpret_cast<float4*>(t.dx + a);
        for (long long i = threadIdx.x; i < ((b - a) >> 2); i += blockDim.x) o[i] = p[i];
    } else {
        for (long long i = a + threadIdx.x; i < b; i += blockDim.x) t.dx[i] = t.dy[i];
    }
}

// ---- mean-based weights: GDRQ_PY with is_weight (core/operator/GDRQ.py:69-74,97-102) -------------------------------
// launch A: one double partial per RANGE item (ROWS items do their own sum in launch B).
__global__ void __launch_bounds__(B2Q_THREADS)
mt_sumabs_kernel(const MtTensor* __restrict__ tensors, const MtItem* __restrict__ items, double* __restrict__ partial) {
    b2q_pdl_sync();
    __shared__ double smem[32];
    const MtItem it = items[blockIdx.x];
    if (it.kind != 0) return;
    const MtTensor t = tensors[it.tensor];
    const float* base = mt_group_base(t, it.group);
    double acc = 0.0;
    for (long long i = it.a + threadIdx.x; i < it.b; i += blockDim.x) acc += (double)fabsf(base[i]);
    const double r = block_reduce<false>(acc, smem);
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
}

// launch B: alpha = ktimes * mean|w| (unless fix_alpha), clip, round to qlevel levels.
__global__ void __launch_bounds__(B2Q_THREADS)
mt_gdrq_kernel(const MtTensor* __restrict__ tensors, const MtItem* __restrict__ items, const double* __restrict__ partial,
               int fix_alpha, int do_round, float qlevel, float ktimes, int fast) {
    b2q_pdl_sync();
    __shared__ double smem[32];
    __shared__ float s_T;
    const MtItem it = items[blockIdx.x];
    const MtTensor t = tensors[it.tensor];
    if (it.kind == 0) {
        if (!fix_alpha) {   // every block of the group combines the group's partials in the same fixed order
            double acc = 0.0;
            for (int i = threadIdx.x; i < it.pcount; i += blockDim.x) acc += partial[it.pfirst + i];
            const double tot = block_reduce<false>(acc, smem);
            if (threadIdx.x == 0) {
                const float count = (float)(t.per_channel ? t.cols : t.rows * t.cols);
                s_T = __fmul_rn(ktimes, __fdiv_rn((float)tot, count));
                if (it.first) t.aux[it.group] = s_T;
            }
        } else if (threadIdx.x == 0) {
            s_T = t.aux[it.group];
        }
        __syncthreads();
        const float T = s_T;
        const QScale s = make_qscale(T, qlevel, fast != 0);
        const int clip = t.per_channel ? B2Q_CLIP_WHERE_LE : B2Q_CLIP_SYM;   // GDRQ.py:109 vs :79
        const long long off = (t.per_channel ? (long long)it.group * t.cols : 0);
        for (long long i = off + it.a + threadIdx.x; i < off + it.b; i += blockDim.x) {
            const float c = clip_value(clip, t.x[i], T);
            t.y[i] = do_round ? __fmul_rn(quant_code(c, s), s.q) : c;
        }
    } else {
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
        for (long long row = it.a + wid; row < it.b; row += nw) {
            const float* p = t.x + row * t.cols;
            float* o = t.y + row * t.cols;
            float T;
            if (!fix_alpha) {
                double acc = 0.0;
                for (long long c = lane; c < t.cols; c += 32) acc += (double)fabsf(p[c]);
                acc = warp_sum(acc);
                T = __fmul_rn(ktimes, __fdiv_rn((float)acc, (float)t.cols));
                if (lane == 0) t.aux[row] = T;
            } else {
                T = t.aux[row];
            }
            const QScale s = make_qscale(T, qlevel, fast != 0);
            for (long long c = lane; c < t.cols; c += 32) {
                const float v = clip_value(B2Q_CLIP_WHERE_LE, p[c], T);
                o[c] = do_round ? __fmul_rn(quant_code(v, s), s.q) : v;
            }
        }
    }
}

extern "C" {

int b2q_multi_plan_create(b2q_ctx* ctx, const b2q_weight_desc* descs, int count, b2q_multi_plan** out) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(descs && out && count >= 1, "bad argument");
    std::vector<MtTensor> tensors;
    std::vector<MtItem> items;
    int gbase = 0, has_grad = 1;
    long long elements = 0;
    for (int k = 0; k < count; ++k) {
        if (!descs[k].dy || !descs[k].dx) has_grad = 0;
        const b2q_weight_desc& d = descs[k];
        B2Q_REQUIRE(d.x && d.y && d.aux && d.rows >= 1 && d.cols >= 1, "bad weight descriptor");
        MtTensor t = {d.x, d.y, d.aux, d.dy, d.dx, (long long)d.rows, (long long)d.cols, d.per_channel ? 1 : 0, gbase};
        const long long n = t.rows * t.cols;
        elements += n;
        if (!t.per_channel) {
            const int first = (int)items.size();
            const int cnt = (int)((n + MT_ITEM_ELEMS - 1) / MT_ITEM_ELEMS);
            for (long long a = 0; a < n; a += MT_ITEM_ELEMS)
                items.push_back({k, 0, 0, a == 0, first, cnt, a, a + MT_ITEM_ELEMS < n ? a + MT_ITEM_ELEMS : n});
            gbase += 1;
        } else if (t.cols <= MT_SHORT_ROW) {
            for (long long r = 0; r < t.rows; r += MT_ROWS_PER_ITEM)
                items.push_back({k, 1, 0, 0, 0, 0, r, r + MT_ROWS_PER_ITEM < t.rows ? r + MT_ROWS_PER_ITEM : t.rows});
            gbase += (int)t.rows;
        } else {
            const int cnt = (int)((t.cols + MT_ITEM_ELEMS - 1) / MT_ITEM_ELEMS);
            for (long long r = 0; r < t.rows; ++r) {
                const int first = (int)items.size();
                for (long long a = 0; a < t.cols; a += MT_ITEM_ELEMS)
                    items.push_back({k, 0, (int)r, a == 0, first, cnt, a,
                                     a + MT_ITEM_ELEMS < t.cols ? a + MT_ITEM_ELEMS : t.cols});
            }
            gbase += (int)t.rows;
        }
        tensors.push_back(t);
    }
    b2q_multi_plan* p = new b2q_multi_plan();
    memset(p, 0, sizeof(*p));
    p->device = ctx->device;
    p->n_tensors = count;
    p->n_items = (int)items.size();
    p->n_groups = gbase;
    p->elements = elements;
    p->has_grad = has_grad;
    cudaError_t e = cudaMalloc(&p->d_tensors, sizeof(MtTensor) * tensors.size());
    if (e == cudaSuccess) e = cudaMalloc(&p->d_items, sizeof(MtItem) * items.size());
    if (e == cudaSuccess) e = cudaMalloc(&p->d_stat, sizeof(unsigned long long) * gbase);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_epoch, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc(&p->d_partial, sizeof(double) * items.size());
    if (e == cudaSuccess) e = cudaMemcpy(p->d_tensors, tensors.data(), sizeof(MtTensor) * tensors.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_items, items.data(), sizeof(MtItem) * items.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(p->d_stat, 0, sizeof(unsigned long long) * gbase);
    if (e == cudaSuccess) e = cudaMemset(p->d_epoch, 0, sizeof(unsigned int));
    if (e != cudaSuccess) {
        cudaFree(p->d_tensors); cudaFree(p->d_items); cudaFree(p->d_stat); cudaFree(p->d_epoch); cudaFree(p->d_partial);
        delete p;
        b2q_set_error(std::string("multi plan allocation failed: ") + cudaGetErrorString(e));
        return 1;
    }
    *out = p;
    return 0;
}

int b2q_multi_plan_destroy(b2q_ctx* ctx, b2q_multi_plan* p) {
    if (!p) return 0;
    if (ctx) cudaSetDevice(ctx->device);
    cudaFree(p->d_tensors); cudaFree(p->d_items); cudaFree(p->d_stat); cudaFree(p->d_epoch); cudaFree(p->d_partial);
    delete p;
    return 0;
}

int b2q_multi_weight_quant_fwd_f32(b2q_ctx* ctx, b2q_multi_plan* p, int variant, int is_train, void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(p && p->device == ctx->device, "plan belongs to another device");
    B2Q_REQUIRE(variant == 0 || variant == 1, "variant must be 0 or 1");
    cudaStream_t st = (cudaStream_t)stream;
    // quant_ops.py:17-31 reduces always and stores aux when training; clip_grad...py:19-36 reduces only when
    // training and otherwise quantises from the stored aux
    const int do_reduce = (variant == 0 || is_train) ? 1 : 0;
    if (do_reduce) {
        b2q_timed_launch tl(ctx, B2Q_KIND_OTHER, 4.0 * (double)p->elements, st);
        b2q_launch(ctx, mt_absmax_kernel, (unsigned)p->n_items, B2Q_THREADS, st, p->d_tensors, p->d_items, p->d_stat, p->d_epoch);
        B2Q_LAUNCH_CHECK(ctx);
    }
    {
        b2q_timed_launch tl(ctx, B2Q_KIND_OTHER, 8.0 * (double)p->elements, st);
        b2q_launch(ctx, mt_qdq_kernel, (unsigned)p->n_items, B2Q_THREADS, st, p->d_tensors, p->d_items, p->d_stat, p->d_epoch,
                   do_reduce, is_train ? 1 : 0, ctx->fast_div);
        B2Q_LAUNCH_CHECK(ctx);
    }
    return 0;
}

int b2q_multi_gdrq_weight_fwd_f32(b2q_ctx* ctx, b2q_multi_plan* p, int fix_alpha, int do_round, float qlevel,
                                  float ktimes, void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(p && p->device == ctx->device, "plan belongs to another device");
    cudaStream_t st = (cudaStream_t)stream;
    if (!fix_alpha) {
        b2q_timed_launch tl(ctx, B2Q_KIND_OTHER, 4.0 * (double)p->elements, st);
        b2q_launch(ctx, mt_sumabs_kernel, (unsigned)p->n_items, B2Q_THREADS, st, p->d_tensors, p->d_items, p->d_partial);
        B2Q_LAUNCH_CHECK(ctx);
    }
    b2q_timed_launch tl(ctx, B2Q_KIND_OTHER, 8.0 * (double)p->elements, st);
    b2q_launch(ctx, mt_gdrq_kernel, (unsigned)p->n_items, B2Q_THREADS, st, p->d_tensors, p->d_items, p->d_partial, fix_alpha,
               do_round, qlevel, ktimes, ctx->fast_div);
    B2Q_LAUNCH_CHECK(ctx);
    return 0;
}

int b2q_multi_weight_ste_bwd_f32(b2q_ctx* ctx, b2q_multi_plan* p, void* stream) {
    B2Q_CTX(ctx);
    B2Q_REQUIRE(p && p->device == ctx->device, "plan belongs to another device");
    B2Q_REQUIRE(p->has_grad, "plan was created without dy/dx pointers");
    cudaStream_t st = (cudaStream_t)stream;
    b2q_timed_launch tl(ctx, B2Q_KIND_OTHER, 8.0 * (double)p->elements, st);
    b2q_launch(ctx, mt_copy_kernel, (unsigned)p->n_items, B2Q_THREADS, st, p->d_tensors, p->d_items);
    B2Q_LAUNCH_CHECK(ctx);
    return 0;
}

}  // extern "C"
