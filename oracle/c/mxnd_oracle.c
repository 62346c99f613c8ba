/* TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- plain-C restatement of the reference operators with
 * MXNet's execution structure: every mx.nd call of the reference op bodies is ONE pass over the tensor with its
 * own float32 temporary (abs -> max -> div -> round -> mul -> assign, symbol/quant_ops.py:34-40), parallelised
 * with OpenMP the way libmxnet's CPU elementwise kernels are.  It is (a) checked bit-for-bit against
 * oracle/quant_oracle.py and (b) the timed "reference CPU path" of bench.py (cpu_baseline / --impl reference),
 * because libmxnet itself cannot be installed in this image.
 *
 * Generosity note for the baseline: libmxnet reduces a whole tensor to one scalar on a single thread
 * (seq_reduce_compute parallelises over OUTPUT elements [upstream]); here full reductions are OpenMP-parallel,
 * Python/GIL dispatch and NDArray allocation are not charged, and temporaries come from a reused workspace.
 *
 * Numerics: roundf (half away), IEEE float32 division, scalar operands applied in float32, sums accumulated
 * in double and rounded once (model of the Kahan-compensated float32 sum), no FMA contraction (-ffp-contract=off).
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    float* a;
    float* b;
    size_t cap;
} ws_t;

static ws_t g_ws = {0, 0, 0};

static int ws_reserve(size_t n) {
    if (n <= g_ws.cap) return 0;
    free(g_ws.a);
    free(g_ws.b);
    g_ws.a = (float*)malloc(n * sizeof(float));
    g_ws.b = (float*)malloc(n * sizeof(float));
    g_ws.cap = (g_ws.a && g_ws.b) ? n : 0;
    return g_ws.cap ? 0 : 1;
}

int b2qo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void b2qo_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- mx.nd primitives: one pass each ----------------------------------------------------------- */
static void nd_abs(const float* x, float* y, size_t n) {
#pragma omp parallel for schedule(static)
    for (ptrdiff_t i = 0; i < (ptrdiff_t)n; ++i) y[i] = fabsf(x[i]);
}

/* NaN propagates (np.max in quant_oracle.py, torch.max in the shim the golden vectors were generated over): the
 * number of NaN elements is counted beside the maximum so the result does not depend on the visiting order */
static float nd_max(const float* x, size_t n) {
    float m = -INFINITY;
    long long nans = 0;
#pragma omp parallel for schedule(static) reduction(max : m) reduction(+ : nans)
    for (ptrdiff_t i = 0; i < (ptrdiff_t)n; ++i) {
        nans += (x[i] != x[i]);
        m = x[i] > m ? x[i] : m;
    }
    return nans ? NAN : m;
}

static float nd_mean(const float* x, size_t n) {
    double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : s)
    for (ptrdiff_t i = 0; i < (ptrdiff_t)n; ++i) s += (double)x[i];
    return (float)s / (float)n;
}

static void nd_div(const float* x, float q, float* y, size_t n) {
#pragma omp parallel for schedule(static)
    for (ptrdiff_t i = 0; i < (ptrdiff_t)n; ++i) y[i] = x[i] / q;
}

static void nd_mul(const float* x, float q, float* y, size_t n) {
#pragma omp parallel for schedule(static)
    for (ptrdiff_t i = 0; i < (ptrdiff_t)n; ++i) y[i] = x[i] * q;
}

static void nd_round(const float* x, float* y, size_t n) {
#pragma omp parallel for schedule(static)
    for (ptrdiff_t i = 0; i < (ptrdiff_t)n; ++i) y[i] = roundf(x[i]);
}

static void nd_clip(const float* x, float lo, float hi, float* y, size_t n) {
#pragma omp parallel for schedule(static)
    for (ptrdiff_t i = 0; i < (ptrdiff_t)n; ++i) {
        float v = x[i];
        y[i] = v > hi ? hi : (v < lo ? lo : v);
    }
}

static void nd_assign(float* dst, const float* src, size_t n, int req) { /* CustomOp.assign */
    if (req == 0) return;
    if (req == 3) {
#pragma omp parallel for schedule(static)
        for (ptrdiff_t i = 0; i < (ptrdiff_t)n; ++i) dst[i] = dst[i] + src[i];
    } else {
#pragma omp parallel for schedule(static)
        for (ptrdiff_t i = 0; i < (ptrdiff_t)n; ++i) dst[i] = src[i];
    }
}

/* y = x * (x cmp t) as the reference writes it: compare kernel -> 0/1 tensor, then multiply kernel */
static void nd_cmp_gt(const float* x, float t, float* y, size_t n) {
#pragma omp parallel for schedule(static)
    for (ptrdiff_t i = 0; i < (ptrdiff_t)n; ++i) y[i] = x[i] > t ? 1.0f : 0.0f;
}
static void nd_cmp_lt(const float* x, float t, float* y, size_t n) {
#pragma omp parallel for schedule(static)
    for (ptrdiff_t i = 0; i < (ptrdiff_t)n; ++i) y[i] = x[i] < t ? 1.0f : 0.0f;
}
static void nd_cmp_le(const float* x, float t, float* y, size_t n) {
#pragma omp parallel for schedule(static)
    for (ptrdiff_t i = 0; i < (ptrdiff_t)n; ++i) y[i] = x[i] <= t ? 1.0f : 0.0f;
}
static void nd_mul_t(const float* a, const float* b, float* y, size_t n) {
#pragma omp parallel for schedule(static)
    for (ptrdiff_t i = 0; i < (ptrdiff_t)n; ++i) y[i] = a[i] * b[i];
}

/* per-row (out-channel) pieces: max over axes 1.., then broadcast_like materialises a full-size scale tensor */
static void nd_max_rows(const float* x, float* out, size_t rows, size_t cols) {
#pragma omp parallel for schedule(static)
    for (ptrdiff_t r = 0; r < (ptrdiff_t)rows; ++r) {
        float m = -INFINITY;
        int nans = 0;
        for (size_t c = 0; c < cols; ++c) {
            nans |= (x[r * cols + c] != x[r * cols + c]);
            m = x[r * cols + c] > m ? x[r * cols + c] : m;
        }
        out[r] = nans ? NAN : m;
    }
}
static void nd_broadcast_rows(const float* v, float* y, size_t rows, size_t cols) {
#pragma omp parallel for schedule(static)
    for (ptrdiff_t r = 0; r < (ptrdiff_t)rows; ++r)
        for (size_t c = 0; c < cols; ++c) y[r * cols + c] = v[r];
}
static void nd_div_t(const float* a, const float* b, float* y, size_t n) {
#pragma omp parallel for schedule(static)
    for (ptrdiff_t i = 0; i < (ptrdiff_t)n; ++i) y[i] = a[i] / b[i];
}

/* round(x / q) * q with q a scalar: three passes, three temporaries (quant_ops.py:28,40) */
static void qdq_scalar(const float* x, float q, float* out, size_t n, int req) {
    nd_div(x, q, g_ws.a, n);
    nd_round(g_ws.a, g_ws.b, n);
    nd_mul(g_ws.b, q, g_ws.a, n);
    nd_assign(out, g_ws.a, n, req);
}

/* ---- operators ----------------------------------------------------------------------------------- */
/* symbol/quant_ops.py:17-40 (variant 0)  /  symbol/clip_grad_quantization_int8.py:19-51 (variant 1) */
int b2qo_minmax_quant_fwd(int variant, const float* x, float* y, float* aux, int64_t rows, int64_t cols,
                          int is_weight, int per_channel, int is_train, int init, float d, float omd, int req) {
    const size_t n = (size_t)rows * (size_t)cols;
    if (ws_reserve(n)) return 1;
    if (is_weight && per_channel) {
        float* maxs = (float*)malloc(sizeof(float) * (size_t)rows);
        float* unit = (float*)malloc(sizeof(float) * (size_t)rows);
        if (!maxs || !unit) return 1;
        if (variant == 0 || is_train) {
            nd_abs(x, g_ws.a, n);
            nd_max_rows(g_ws.a, maxs, (size_t)rows, (size_t)cols);
            if (is_train) memcpy(aux, maxs, sizeof(float) * (size_t)rows);
        }
        const float* src = (variant == 0) ? maxs : aux;
        for (int64_t r = 0; r < rows; ++r) unit[r] = src[r] / 127.0f;
        nd_broadcast_rows(unit, g_ws.b, (size_t)rows, (size_t)cols);
        nd_div_t(x, g_ws.b, g_ws.a, n);
        nd_round(g_ws.a, g_ws.a, n);
        nd_mul_t(g_ws.a, g_ws.b, g_ws.a, n);
        nd_assign(y, g_ws.a, n, req);
        free(maxs);
        free(unit);
        return 0;
    }
    if (is_weight) {
        float m = aux[0];
        if (variant == 0 || is_train) {
            nd_abs(x, g_ws.a, n);
            m = nd_max(g_ws.a, n);
            if (is_train) aux[0] = m;
        }
        if (variant == 1) m = aux[0];
        qdq_scalar(x, m / 127.0f, y, n, req);
        return 0;
    }
    if (is_train) {
        nd_abs(x, g_ws.a, n);
        const float m = nd_max(g_ws.a, n);
        aux[0] = (variant == 1 && init) ? m : (aux[0] * d + m * omd);
    }
    const float q = aux[0] / 127.0f;
    if (variant == 1) { /* clip to +-aux, written with [:]= (ignores req) */
        nd_clip(x, -aux[0], aux[0], y, n);
        nd_div(y, q, g_ws.a, n);
        nd_round(g_ws.a, g_ws.b, n);
        nd_mul(g_ws.b, q, g_ws.a, n);
        nd_assign(y, g_ws.a, n, 1);
    } else {
        qdq_scalar(x, q, y, n, req);
    }
    return 0;
}

/* quant_ops.py:41-42 */
int b2qo_ste_bwd(const float* dy, float* dx, int64_t n, int req) {
    nd_assign(dx, dy, (size_t)n, req);
    return 0;
}

/* clip_grad_quantization_int8.py:61-67: copy, compare, multiply, copy, compare, multiply, copy */
int b2qo_clipgrad_bwd(const float* x, const float* dy, float* dx, const float* aux, int64_t n64) {
    const size_t n = (size_t)n64;
    if (ws_reserve(n)) return 1;
    nd_assign(dx, dy, n, 1);
    nd_cmp_gt(x, -aux[0], g_ws.a, n);
    nd_mul_t(dx, g_ws.a, g_ws.b, n);
    nd_assign(dx, g_ws.b, n, 1);
    nd_cmp_lt(x, aux[0], g_ws.a, n);
    nd_mul_t(dx, g_ws.a, g_ws.b, n);
    nd_assign(dx, g_ws.b, n, 1);
    return 0;
}

/* core/operator/GDRQ.py:69-86 (group_size == -1) */
int b2qo_gdrq_fwd(const float* x, float* y, float* alpha, int64_t n64, int is_weight, int fix_alpha, int do_round,
                  float qlevel, float ktimes, float lamda, int req) {
    const size_t n = (size_t)n64;
    if (ws_reserve(n)) return 1;
    nd_abs(x, g_ws.a, n);                      /* GDRQ.py:67 (always computed) */
    if (!fix_alpha) {
        const float thr = ktimes * nd_mean(g_ws.a, n);
        if (is_weight) alpha[0] = thr;
        else {
            const float diff = alpha[0] - thr;
            const float step = lamda * diff;
            alpha[0] = alpha[0] + step;
        }
    }
    nd_clip(x, -alpha[0], alpha[0], g_ws.b, n);
    if (do_round) {
        const float q = alpha[0] / qlevel;
        nd_div(g_ws.b, q, g_ws.a, n);
        nd_round(g_ws.a, g_ws.b, n);
        nd_mul(g_ws.b, q, g_ws.a, n);
        nd_assign(y, g_ws.a, n, req);
    } else {
        nd_assign(y, g_ws.b, n, req);
    }
    return 0;
}

/* GDRQ.py:131-133 */
int b2qo_gdrq_bwd(const float* x, const float* dy, float* dx, const float* alpha, int64_t n64, int req) {
    const size_t n = (size_t)n64;
    if (ws_reserve(n)) return 1;
    nd_abs(x, g_ws.a, n);
    nd_cmp_le(g_ws.a, alpha[0], g_ws.b, n);
    nd_mul_t(dy, g_ws.b, g_ws.a, n);
    nd_assign(dx, g_ws.a, n, req);
    return 0;
}

/* symbol/fold_bn_v1_gdrq.py:53-68 (data side, training) */
int b2qo_foldbn_data_fwd(const float* x, float* y, float* aux, int64_t n64, int init, float d, float omd) {
    const size_t n = (size_t)n64;
    if (ws_reserve(n)) return 1;
    nd_abs(x, g_ws.a, n);
    const float thr = 2.0f * nd_mean(g_ws.a, n);
    aux[0] = init ? thr : (aux[0] * d + thr * omd);
    const float q = aux[0] / 127.0f;
    nd_clip(x, -thr, thr, g_ws.b, n);
    nd_div(g_ws.b, q, g_ws.a, n);
    nd_round(g_ws.a, g_ws.b, n);
    nd_mul(g_ws.b, q, y, n);
    return 0;
}

/* symbol/fold_bn_v1_gdrq.py:70-96,113 (weight side) */
int b2qo_foldbn_weight_fwd(const float* w, float* wq, float* bias, float* aux, const float* gamma, const float* beta,
                           const float* mean, const float* var, float eps, int64_t cout, int64_t cols, int per_channel,
                           int quantize, int is_train) {
    const size_t n = (size_t)cout * (size_t)cols;
    if (ws_reserve(n)) return 1;
    float* factor = (float*)malloc(sizeof(float) * (size_t)cout);
    if (!factor) return 1;
    for (int64_t c = 0; c < cout; ++c) {
        const float den = sqrtf(var[c] + eps);
        factor[c] = gamma[c] / den;
        const float prod = mean[c] * gamma[c];
        bias[c] = beta[c] - prod / den;
    }
#pragma omp parallel for schedule(static)
    for (ptrdiff_t r = 0; r < (ptrdiff_t)cout; ++r)
        for (int64_t c = 0; c < cols; ++c) wq[r * cols + c] = w[r * cols + c] * factor[r];
    free(factor);
    if (!quantize) return 0;
    nd_abs(wq, g_ws.a, n);
    if (per_channel) {
        for (int64_t r = 0; r < cout; ++r) {   /* python loop over channels in the reference (:84-85) */
            const float thr = 2.0f * nd_mean(g_ws.a + r * cols, (size_t)cols);
            const float q = thr / 127.0f;
            if (is_train) aux[r] = thr;
            float* row = wq + r * cols;
            for (int64_t c = 0; c < cols; ++c) {
                float v = row[c];
                v = v > thr ? thr : (v < -thr ? -thr : v);
                const float t = v / q;
                row[c] = roundf(t) * q;
            }
        }
    } else {
        const float thr = 2.0f * nd_mean(g_ws.a, n);
        const float q = thr / 127.0f;
        if (is_train) aux[0] = thr;
        nd_clip(wq, -thr, thr, g_ws.b, n);
        nd_div(g_ws.b, q, g_ws.a, n);
        nd_round(g_ws.a, g_ws.b, n);
        nd_mul(g_ws.b, q, wq, n);
    }
    return 0;
}
