"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- NumPy restatement of the reference fake-quant ops.

Each class restates one reference ``mx.operator.CustomOp`` with the *same* call protocol
(``forward(is_train, req, in_data, out_data, aux)`` / ``backward(req, out_grad, in_data, out_data,
in_grad, aux)``; lists of float32 ``numpy.ndarray`` mutated in place), so a test drives the CUDA op
and the oracle identically.  Citations are ``/root/reference/<file>:<line>``.

MXNet numerics encoded here ([upstream] assumptions, SURVEY.md section 8c -- libmxnet itself cannot
be installed in this image, so these are stated, not measured):

* every ``mx.nd`` call / NDArray operator is its own kernel with a float32 result -> ``fl()`` after
  every step, never a fused multiply-add;
* ``mx.nd.round`` is C ``roundf`` = half away from zero (``mx_round``); NOT ``np.round``;
* ``/`` is IEEE float32 division;
* Python scalars (``127``, ``ema_decay``, ``1 - ema_decay``, ``2``, ``ktimes``, ``lamda``, ``eps``) are
  computed in Python double and cast to float32 when applied;
* ``mx.nd.clip(x, lo, hi)`` is ``x > hi ? hi : (x < lo ? lo : x)`` (NaN passes through);
* ``mx.nd.sign(0) = 0``; comparisons give 0.0 / 1.0 float32;
* ``mx.nd.max`` is exact; ``mx.nd.mean`` = float32 sum (Kahan-compensated in libmxnet; modelled here as
  the correctly rounded sum, i.e. float64 accumulation rounded once) divided by float32(N);
* a ``(1,)`` aux broadcasts against N-d data.

The oracle also exposes the integer "codes" (``round(x/q)`` before re-scaling) that BASELINE.json
requires to be bit-exact.
"""
import numpy as np

F = np.float32
_ERR = dict(divide="ignore", invalid="ignore", over="ignore", under="ignore")


# ----------------------------------------------------------------------------------------------
# mx.nd primitives
# ----------------------------------------------------------------------------------------------
def fl(x):
    return np.asarray(x, dtype=F)


def mx_round(x):
    """roundf: half away from zero, exact (not floor(|x|+0.5), which is wrong at 0.49999997)."""
    x = fl(x)
    t = np.trunc(x)
    with np.errstate(**_ERR):
        r = x - t  # exact
    step = np.where(np.abs(r) >= F(0.5), np.sign(x), F(0)).astype(F)
    out = (t + step).astype(F)
    # keep the sign of zero the way roundf does (roundf(-0.3) == -0.0)
    return np.copysign(out, x).astype(F)


def mx_div(a, b):
    with np.errstate(**_ERR):
        return (fl(a) / fl(b)).astype(F)


def mx_mul(a, b):
    with np.errstate(**_ERR):
        return (fl(a) * fl(b)).astype(F)


def mx_add(a, b):
    with np.errstate(**_ERR):
        return (fl(a) + fl(b)).astype(F)


def mx_sub(a, b):
    with np.errstate(**_ERR):
        return (fl(a) - fl(b)).astype(F)


def mx_sign(x):
    """mshadow_op::sign [upstream]: a < 0 ? -1 : (a > 0 ? 1 : 0) -- so sign(0) = 0 and sign(NaN) = 0 (np.sign gives NaN)."""
    x = fl(x)
    return np.where(x > 0, F(1), np.where(x < 0, F(-1), F(0))).astype(F)


def mx_clip(x, lo, hi):
    x = fl(x)
    lo = F(lo)
    hi = F(hi)
    return np.where(x > hi, hi, np.where(x < lo, lo, x)).astype(F)


def mx_sum(x, axis=None):
    return np.sum(np.asarray(x, dtype=np.float64), axis=axis).astype(F)


def mx_mean(x, axis=None):
    x = np.asarray(x)
    s = mx_sum(x, axis=axis)
    n = x.size // max(int(np.size(s)), 1)
    return mx_div(s, F(n))


def mx_absmax(x, axis=None):
    return np.max(np.abs(fl(x)), axis=axis).astype(F)


def _rest_axes(x):
    return tuple(range(1, x.ndim))


def _col(v, x):
    """reshape a per-channel vector (C,) to (C,1,1,..) against x."""
    return fl(v).reshape((x.shape[0],) + (1,) * (x.ndim - 1))


def qdq(x, q):
    """fl(roundf(fl(x/q)) * q) and the integer-valued codes."""
    codes = mx_round(mx_div(x, q))
    return mx_mul(codes, q), codes


def clip_by_mode(mode, x, t):
    """The six clipping expressions of the reference, numbered like include/b2q.h's B2Q_CLIP_*:
    0 none; 1 mx.nd.clip(x, -t, t) (quant clip_grad...py:48, GDRQ.py:79, fold_bn_v1_gdrq.py:67);
    2 where(|x| <= t, x, t*sign(x)) (GDRQ.py:109); 3 mx.nd.clip(x, 0, t) (GDRQ.py:202);
    4 where(x < t, x, t) (PACT.py:125); 5 where(|x| < t, x, t*sign(x)) (PACT.py:193)."""
    x = fl(x)
    t = F(t)
    if mode == 0:
        return x
    if mode == 1:
        return mx_clip(x, -t, t)
    if mode == 2:
        return np.where(np.abs(x) <= t, x, mx_mul(t, mx_sign(x))).astype(F)
    if mode == 3:
        return mx_clip(x, F(0), t)
    if mode == 4:
        return np.where(x < t, x, t).astype(F)
    if mode == 5:
        return np.where(np.abs(x) < t, x, mx_mul(t, mx_sign(x))).astype(F)
    raise ValueError(mode)


def assign(dst, req, src):
    """mx.operator.CustomOp.assign [upstream python/mxnet/operator.py]."""
    if req == "null":
        return
    if req in ("write", "inplace"):
        dst[...] = src
    elif req == "add":
        dst[...] = mx_add(dst, src)
    else:
        raise ValueError("unknown req %r" % (req,))


class _Op(object):
    codes = None  # last integer codes (test side-channel)

    def assign(self, dst, req, src):
        assign(dst, req, src)


# ----------------------------------------------------------------------------------------------
# symbol/quant_ops.py:3-42   op_type "Quantization_int8_V2"
# ----------------------------------------------------------------------------------------------
class Quantization_int8(_Op):
    def __init__(self, quant_mode, is_weight, is_weight_perchannel, delay_quant, ema_decay):
        self.quant_mode = quant_mode
        self.is_weight = is_weight
        self.is_weight_perchannel = is_weight_perchannel
        self.delay_quant = delay_quant
        self.ema_decay = ema_decay
        self.QUANT_LEVEL = 127  # quant_ops.py:10
        self.init = True        # quant_ops.py:11 (never read)

    def forward(self, is_train, req, in_data, out_data, aux):
        x = in_data[0]
        if is_train and self.delay_quant > 0:           # quant_ops.py:13-16
            self.assign(out_data[0], req[0], x)
            self.delay_quant -= 1
            return
        if self.is_weight:                               # quant_ops.py:17-31
            if self.is_weight_perchannel:
                maxs = mx_absmax(x, axis=_rest_axes(x))  # :20-22
                quant_unit = _col(mx_div(maxs, F(self.QUANT_LEVEL)), x)  # :23-24
            else:
                maxs = mx_absmax(x)                      # :26
                quant_unit = mx_div(maxs, F(self.QUANT_LEVEL))           # :27
            y, self.codes = qdq(x, quant_unit)           # :28
            self.assign(out_data[0], req[0], y)
            if is_train:
                aux[0][...] = maxs                       # :30-31
        else:                                            # quant_ops.py:32-40
            if is_train:
                maxs = mx_absmax(x)                      # :34-35
                aux[0][...] = mx_add(mx_mul(aux[0], F(self.ema_decay)),
                                     mx_mul(maxs, F(1 - self.ema_decay)))  # :37
            quant_unit = mx_div(aux[0], F(self.QUANT_LEVEL))              # :39
            y, self.codes = qdq(x, quant_unit)           # :40  (no clip)
            self.assign(out_data[0], req[0], y)

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        self.assign(in_grad[0], req[0], out_grad[0])     # quant_ops.py:41-42


# ----------------------------------------------------------------------------------------------
# symbol/clip_grad_quantization_int8.py:5-67   op_type "ClipGrad_Quantization_int8"
# ----------------------------------------------------------------------------------------------
class ClipGrad_Quantization_int8(_Op):
    def __init__(self, quant_mode, is_weight, is_weight_perchannel, delay_quant, ema_decay):
        self.quant_mode = quant_mode
        self.is_weight = is_weight
        self.is_weight_perchannel = is_weight_perchannel
        self.delay_quant = delay_quant
        self.ema_decay = ema_decay
        self.QUANT_LEVEL = 127
        self.init = True

    def forward(self, is_train, req, in_data, out_data, aux):
        x = in_data[0]
        if is_train and self.delay_quant > 0:            # clip_grad...py:15-18
            self.assign(out_data[0], req[0], x)
            self.delay_quant -= 1
            return
        if self.is_weight:                               # :19-36
            if self.is_weight_perchannel:
                if is_train > 0:
                    aux[0][...] = mx_absmax(x, axis=_rest_axes(x))       # :22-26
                quant_unit = _col(mx_div(aux[0], F(self.QUANT_LEVEL)), x)  # :27-29
            else:
                if is_train > 0:
                    aux[0][...] = mx_absmax(x)           # :31-34
                quant_unit = mx_div(aux[0], F(self.QUANT_LEVEL))         # :35
            y, self.codes = qdq(x, quant_unit)           # :36
            self.assign(out_data[0], req[0], y)
        else:                                            # :37-51
            if is_train:
                maxs = mx_absmax(x)                      # :39-40
                if self.init:                            # :42-44
                    aux[0][...] = maxs
                    self.init = False
                else:                                    # :46
                    aux[0][...] = mx_add(mx_mul(aux[0], F(self.ema_decay)),
                                         mx_mul(maxs, F(1 - self.ema_decay)))
            quant_unit = mx_div(aux[0], F(self.QUANT_LEVEL))             # :47
            t = float(aux[0][0])                         # :49-50 asnumpy()[0]
            clipped = mx_clip(x, -t, t)                  # :48  written with [:]= (ignores req)
            y, self.codes = qdq(clipped, quant_unit)     # :51
            out_data[0][...] = y

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        if self.is_weight:                               # :58-59
            self.assign(in_grad[0], req[0], out_grad[0])
        else:                                            # :61-67  ([:]= ignores req)
            t = aux[0][0]
            g = fl(out_grad[0]).copy()
            g = mx_mul(g, (in_data[0] > -t).astype(F))
            g = mx_mul(g, (in_data[0] < t).astype(F))
            in_grad[0][...] = g


# ----------------------------------------------------------------------------------------------
# symbol/fold_bn_v1_gdrq.py:5-129   op_type "GDRQ_Fold_BN"
# ----------------------------------------------------------------------------------------------
def conv2d_nchw(x, w, stride, pad, dilate, num_group):
    """Plain direct convolution (oracle for mx.nd.Convolution, fold_bn_v1_gdrq.py:99-110); float64
    accumulation rounded once -- the conv itself is library code (cuDNN) on both sides, so it is only
    compared to tolerance."""
    x = np.asarray(x, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    n, c, h, wd = x.shape
    co, cig, kh, kw = w.shape
    sh, sw = stride
    ph, pw = pad
    dh, dw = dilate
    g = num_group
    assert c == cig * g and co % g == 0
    oh = (h + 2 * ph - (dh * (kh - 1) + 1)) // sh + 1
    ow = (wd + 2 * pw - (dw * (kw - 1) + 1)) // sw + 1
    xp = np.zeros((n, c, h + 2 * ph, wd + 2 * pw), dtype=np.float64)
    xp[:, :, ph:ph + h, pw:pw + wd] = x
    y = np.zeros((n, co, oh, ow), dtype=np.float64)
    cog = co // g
    for gi in range(g):
        xs = xp[:, gi * cig:(gi + 1) * cig]
        ws = w[gi * cog:(gi + 1) * cog]
        for i in range(kh):
            for j in range(kw):
                patch = xs[:, :, i * dh:i * dh + sh * (oh - 1) + 1:sh, j * dw:j * dw + sw * (ow - 1) + 1:sw]
                y[:, gi * cog:(gi + 1) * cog] += np.einsum("nchw,oc->nohw", patch, ws[:, :, i, j])
    return y.astype(F)


def bn_v1_batch_stats(x):
    """Batch statistics BatchNorm_v1 emits with output_mean_var=True -- what feeds bn_mean / bn_var of GDRQ_Fold_BN in
    the reference's graph (/root/reference/symbol/fold_bn_v1_gdrq.py:271-272).
    [upstream src/operator/batch_norm_v1-inl.h]  scale = fl(C / size);  mean = scale * sumall_except_dim<1>(data);
    var = scale * sumall_except_dim<1>(square(data - broadcast<1>(mean))): every element-wise step float32, the two
    reductions modelled as correctly rounded sums (float64 accumulation rounded once), biased variance."""
    x = fl(x)
    c = x.shape[1]
    axes = (0,) + tuple(range(2, x.ndim))
    scale = F(F(c) / F(x.size))
    mean = mx_mul(scale, mx_sum(x, axis=axes))
    dev = mx_sub(x, mean.reshape((1, c) + (1,) * (x.ndim - 2)))
    var = mx_mul(scale, mx_sum(mx_mul(dev, dev), axis=axes))
    return mean, var


class GDRQ_Fold_BN(_Op):
    def __init__(self, quant_mode, is_weight_perchannel, delay_quant, ema_decay,
                 name, num_filter, num_group, kernel, stride, pad, dilate, no_bias,
                 eps, momentum, fix_gamma, quantize_flag):
        self.quant_mode = quant_mode
        self.is_weight_perchannel = is_weight_perchannel
        self.delay_quant = delay_quant
        self.ema_decay = ema_decay
        self.QUANT_LEVEL = 127
        self.init = True
        self.name = name
        self.num_filter = num_filter
        self.num_group = num_group
        self.kernel = kernel
        self.stride = stride
        self.pad = pad
        self.dilate = dilate
        self.no_bias = no_bias
        assert self.no_bias == True, "fold bn don't support bias mode in conv or deconv"  # :24
        self.eps = eps
        self.momentum = momentum
        self.fix_gamma = fix_gamma
        self.quantize_flag = quantize_flag
        # test side-channels
        self.data_q = None
        self.weight_q = None
        self.bias = None
        self.data_codes = None
        self.weight_codes = None

    def forward(self, is_train, req, in_data, out_data, aux):
        assert len(in_data) == 7                                           # :33
        data, weight, bn_output, bn_gamma, bn_beta, bn_mean, bn_var = in_data
        if is_train and self.delay_quant > 0:                              # :43-47
            self.assign(out_data[0], req[0], bn_output)
            self.delay_quant -= 1
            return
        data = fl(data)
        if self.quantize_flag:                                             # :53-68
            if is_train:
                thresholds = mx_mul(F(2), mx_mean(np.abs(data)))           # :56-58
                if self.init:                                              # :60-62
                    aux[0][...] = thresholds
                    self.init = False
                else:                                                      # :64
                    aux[0][...] = mx_add(mx_mul(aux[0], F(self.ema_decay)),
                                         mx_mul(thresholds, F(1 - self.ema_decay)))
            quant_unit = mx_div(aux[0], F(self.QUANT_LEVEL))               # :65
            # :67 -- `thresholds` is only bound when is_train: the reference raises NameError otherwise
            if not is_train:
                raise NameError("name 'thresholds' is not defined")
            t = float(np.reshape(thresholds, -1)[0])
            data = mx_clip(data, -t, t)                                    # :67 (batch threshold)
            data, self.data_codes = qdq(data, quant_unit)                  # :68 (EMA scale)
        self.data_q = data

        factor = mx_div(bn_gamma, np.sqrt(mx_add(bn_var, F(self.eps))).astype(F))  # :72
        weight = mx_mul(weight, _col(factor, weight))                      # :73-74

        if self.quantize_flag:                                             # :76-96
            wabs = np.abs(weight)
            if self.is_weight_perchannel:
                thresholds = mx_mul(F(2), mx_mean(wabs, axis=_rest_axes(weight)))  # :82
                weight = weight.copy()
                for i in range(weight.shape[0]):                           # :84-85
                    weight[i] = mx_clip(weight[i], -float(thresholds[i]), float(thresholds[i]))
                quant_unit = _col(mx_div(thresholds, F(self.QUANT_LEVEL)), weight)  # :86-87
            else:
                thresholds = mx_mul(F(2), mx_mean(wabs))                   # :90
                t = float(thresholds)
                weight = mx_clip(weight, -t, t)                            # :92
                quant_unit = mx_div(thresholds, F(self.QUANT_LEVEL))       # :93
            if is_train:
                aux[1][...] = thresholds                                   # :94-95
            weight, self.weight_codes = qdq(weight, quant_unit)            # :96
        self.weight_q = weight

        conv = conv2d_nchw(data, weight, self.stride, self.pad, self.dilate, self.num_group)  # :99-110
        bias = mx_sub(bn_beta, mx_div(mx_mul(bn_mean, bn_gamma),
                                      np.sqrt(mx_add(bn_var, F(self.eps))).astype(F)))  # :113
        self.bias = bias
        self.assign(out_data[0], req[0], mx_add(conv, bias.reshape(1, -1, 1, 1)))   # :115-120

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        for i in range(len(in_data)):                                      # :124-125
            in_grad[i][...] = 0
        self.assign(in_grad[2], req[2], out_grad[0])                       # :129


# ----------------------------------------------------------------------------------------------
# core/operator/GDRQ.py:50-152   op_type "GDRQ_PY"
# ----------------------------------------------------------------------------------------------
class GDRQ_PY(_Op):
    def __init__(self, nbits, group_size, is_weight, lamda, delay_quant, fix_alpha, ktimes):
        self.nbits = nbits
        self.group_size = group_size
        self.is_weight = is_weight
        self.lamda = lamda
        self.delay_quant = delay_quant
        self.QUANT_LEVEL = 2 ** (self.nbits) - 1                           # GDRQ.py:57
        self.fix_alpha = fix_alpha
        self.ktimes = ktimes

    def _alpha_update(self, alpha, threshold):
        if self.is_weight:
            alpha[...] = threshold                                         # :74 / :102
        else:                                                              # :76 / :104
            alpha[...] = mx_add(alpha, mx_mul(F(self.lamda), mx_sub(alpha, threshold)))

    def _grouped_view(self, a):
        if self.is_weight is False:
            a = np.swapaxes(a, 0, 1)                                       # :90
        shape = a.shape
        return a.reshape((shape[0] // self.group_size, self.group_size) + shape[1:]), shape  # :92-94

    def forward(self, is_train, req, in_data, out_data, aux):
        data = fl(in_data[0])
        alpha = aux[0]
        if self.group_size == -1:                                          # :69-86
            if self.fix_alpha is False:
                threshold = mx_mul(F(self.ktimes), mx_mean(np.abs(data)))  # :71-72
                self._alpha_update(alpha, threshold)
            t = float(alpha[0])                                            # :78
            clipped = mx_clip(data, -t, t)                                 # :79
            if self.delay_quant > 0:                                       # :81-82
                self.delay_quant -= 1
                self.codes = None
            else:                                                          # :84-85
                quant_unit = mx_div(alpha, F(self.QUANT_LEVEL))
                clipped, self.codes = qdq(clipped, quant_unit)
            self.assign(out_data[0], req[0], clipped)                      # :86
        else:                                                              # :88-118
            r, shape = self._grouped_view(data)
            rabs = np.abs(r)
            rsign = mx_sign(r)
            if self.fix_alpha is False:
                threshold = mx_mul(F(self.ktimes), mx_mean(rabs, axis=tuple(range(1, r.ndim))))  # :98-100
                self._alpha_update(alpha, threshold)
            ra = fl(alpha).reshape(alpha.shape + (1,) * (r.ndim - 1))      # :106-107
            clipped = np.where(rabs <= ra, r, mx_mul(ra, rsign)).astype(F)  # :109
            if self.delay_quant > 0:
                self.delay_quant -= 1
                self.codes = None
            else:                                                          # :113-114
                quant_unit = mx_div(ra, F(self.QUANT_LEVEL))
                clipped, codes = qdq(clipped, quant_unit)
                codes = codes.reshape(shape)
                self.codes = codes if self.is_weight else np.swapaxes(codes, 0, 1)
            clipped = clipped.reshape(shape)                               # :115
            if self.is_weight is False:
                clipped = np.swapaxes(clipped, 0, 1)                       # :117
            self.assign(out_data[0], req[0], clipped)                      # :118

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        if self.is_weight:                                                 # :125-127
            self.assign(in_grad[0], req[0], out_grad[0])
            return
        data = fl(in_data[0])
        alpha = aux[0]
        if self.group_size == -1:                                          # :131-133
            flag = (np.abs(data) <= alpha).astype(F)
            self.assign(in_grad[0], req[0], mx_mul(out_grad[0], flag))
        else:                                                              # :134-152
            r, shape = self._grouped_view(data)
            g, _ = self._grouped_view(fl(out_grad[0]))
            ra = fl(alpha).reshape(alpha.shape + (1,) * (r.ndim - 1))
            flag = (np.abs(r) <= ra).astype(F)
            cg = mx_mul(g, flag).reshape(shape)
            if self.is_weight is False:
                cg = np.swapaxes(cg, 0, 1)
            self.assign(in_grad[0], req[0], cg)


# core/operator/GDRQ.py:192-208   op_type "CLIP_RELU_PY"
class CLIP_RELU_PY(_Op):
    def __init__(self, nbits, threshold):
        self.nbits = nbits
        self.threshold = threshold
        self.QUANT_LEVEL = 2 ** (self.nbits) - 1

    def forward(self, is_train, req, in_data, out_data, aux):
        clipped = mx_clip(in_data[0], 0, self.threshold)                   # :202
        quant_unit = F(self.threshold / self.QUANT_LEVEL)                  # :203 (python double, then f32)
        y, self.codes = qdq(clipped, quant_unit)                           # :204
        self.assign(out_data[0], req[0], y)

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        flag = (fl(in_data[0]) < F(self.threshold)).astype(F)              # :207
        self.assign(in_grad[0], req[0], mx_mul(out_grad[0], flag))         # :208


# ----------------------------------------------------------------------------------------------
# core/operator/PACT.py:238-252   op_type "QUANT_STE_PY"
# ----------------------------------------------------------------------------------------------
class QUANT_STE_PY(_Op):
    def __init__(self, nbits):
        self.nbits = nbits
        self.QUANT_LEVEL = 2 ** (self.nbits - 1) - 1                       # PACT.py:241

    def forward(self, is_train, req, in_data, out_data, aux):
        x = in_data[0]
        quant_unit = mx_div(mx_absmax(x), F(self.QUANT_LEVEL))             # :247-248
        y, self.codes = qdq(x, quant_unit)                                 # :249
        self.assign(out_data[0], req[0], y)

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        self.assign(in_grad[0], req[0], out_grad[0])                       # :252


# core/operator/PACT.py:102-144   op_type "PACT_PY";  :169-203 "PACT_V2_PY"
class PACT_PY(_Op):
    two_sided = False

    def __init__(self, nbits):
        self.nbits = nbits
        self.QUANT_LEVEL = 2 ** self.nbits - 1

    def _clip(self, x, gamma):
        if self.two_sided:                                                 # :191-193
            return np.abs(x) < gamma, mx_mul(gamma, mx_sign(x))
        return x < gamma, np.broadcast_to(gamma, x.shape)                  # :125

    def forward(self, is_train, req, in_data, out_data, aux):
        assert len(in_data) == 2
        x, gamma = fl(in_data[0]), fl(in_data[1])
        cond, other = self._clip(x, gamma)
        output = np.where(cond, x, other).astype(F)
        quant_unit = mx_div(gamma, F(self.QUANT_LEVEL))                    # :127 / :197
        y, self.codes = qdq(output, quant_unit)                            # :128 / :198
        self.assign(out_data[0], req[0], y)

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        # autograd of mx.nd.where (PACT.py:142-144 / :201-203): the head gradient goes to the selected branch
        x, gamma = fl(in_data[0]), fl(in_data[1])
        dy = fl(out_grad[0])
        cond, _ = self._clip(x, gamma)
        dx = np.where(cond, dy, F(0)).astype(F)
        dother = np.where(cond, F(0), dy).astype(F)
        if self.two_sided:  # d(gamma_b * sign(x)) / d gamma_b = sign(x)
            dother = mx_mul(dother, mx_sign(x))
        dgamma = mx_sum(dother).reshape(1)
        self.assign(in_grad[0], req[0], dx)
        self.assign(in_grad[1], req[1], dgamma)


class PACT_V2_PY(PACT_PY):
    two_sided = True


# core/operator/PACT.py:30-77   op_type "DoReFa_PY"
class DoReFa_PY(_Op):
    def __init__(self, nbits):
        self.nbits = nbits

    def forward(self, is_train, req, in_data, out_data, aux):
        L = F(2 ** self.nbits - 1)                                         # quantizeK, PACT.py:26-28
        t = np.tanh(fl(in_data[0])).astype(F)                              # :47
        v = mx_absmax(t)                                                   # :48
        o = mx_add(mx_div(t, mx_mul(F(2), v)), F(0.5))                     # :49
        self._t, self._v = t, v
        self.codes = mx_round(mx_mul(L, o))
        y = mx_sub(mx_mul(F(2), mx_div(self.codes, L)), F(1))              # :50
        self.assign(out_data[0], req[0], y)

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        # autograd of  o = t/(2 v) + 0.5,  v = max|t|,  t = tanh(x)  seeded with 2*dy  (PACT.py:76-77).
        # max backward sends the gradient to every element equal to the max [upstream]; abs backward is sign().
        t, v = self._t, self._v
        g = mx_mul(F(2), out_grad[0])
        two_v = mx_mul(F(2), v)
        dt = mx_div(g, two_v)
        # d o / d (2v) = -t / (2v)^2 ; d(2v)/dv = 2
        d2v = mx_sum(mx_mul(g, mx_div(-t, mx_mul(two_v, two_v))))
        dv = mx_mul(F(2), d2v)
        dt = mx_add(dt, mx_mul(mx_mul((np.abs(t) == v).astype(F), dv), mx_sign(t)))
        dx = mx_mul(dt, mx_sub(F(1), mx_mul(t, t)))
        self.assign(in_grad[0], req[0], dx)


# ----------------------------------------------------------------------------------------------
# core/operator/WNQ.py:46-85   op_type "WNQ_PY"
# ----------------------------------------------------------------------------------------------
class WNQ_PY(_Op):
    def __init__(self, nbits, is_perchannel):
        self.nbits = nbits
        self.is_perchannel = is_perchannel
        self.QUANT_LEVEL = 2 ** self.nbits - 1

    def _max(self, x):
        if self.is_perchannel is False:
            return mx_absmax(x)                                            # WNQ.py:55
        return _col(mx_absmax(x, axis=_rest_axes(x)), x)                   # :57-60

    def forward(self, is_train, req, in_data, out_data, aux):
        x = fl(in_data[0])
        m = self._max(x)
        L = F(self.QUANT_LEVEL)
        normed = mx_div(x, m)                                              # :62
        self.codes = mx_round(mx_mul(normed, L))
        self.assign(out_data[0], req[0], mx_mul(mx_div(self.codes, L), m))  # :63

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        x = fl(in_data[0])
        dy = fl(out_grad[0])
        xabs = np.abs(x)
        m = self._max(x)
        not_max = (xabs != m).astype(F)                                    # :71 / :80
        is_max = (xabs == m).astype(F)                                     # :72 / :81
        prod = mx_mul(mx_mul(dy, x), not_max)
        if self.is_perchannel is False:
            mgrad = mx_div(-mx_sum(prod), m)                               # :73
        else:
            mgrad = mx_div(-_col(mx_sum(prod, axis=_rest_axes(x)), x), m)  # :83
        self.assign(in_grad[0], req[0], mx_add(mx_mul(dy, not_max), mx_mul(mgrad, is_max)))  # :85


# ----------------------------------------------------------------------------------------------
# core/operator/QIL.py:34-124, QIL_V2.py:18-71, QIL_V3.py:21-70   op_types "QIL_PY" "QIL_V2_PY" "QIL_V3_PY"
# ----------------------------------------------------------------------------------------------
class _QILBase(_Op):
    """The three QIL variants share  out = sign(x) * ([|x| > cp] + (a|x| + b) [pp <= |x| <= cp]),
    rounded to L levels; they differ in how (pp, cp, a, b) derive from the two learnable scalars and
    therefore in the scalar gradients.  Backward restates the autograd graph analytically."""

    def __init__(self, is_weight, fix_gamma, nbits):
        self.is_weight = is_weight
        self.fix_gamma = fix_gamma
        self.nbits = nbits
        self.QUANT_LEVEL = 2 ** self.nbits - 1

    def _transform(self, x, pp, cp, a, b):
        xabs = np.abs(x)
        sgn = mx_sign(x)
        inter = ((xabs >= pp).astype(F) * (xabs <= cp).astype(F)).astype(F)
        lin = mx_add(mx_mul(a, xabs), b)
        out = mx_add(mx_mul(sgn, (xabs > cp).astype(F)), mx_mul(mx_mul(sgn, lin), inter))
        return out, xabs, sgn, inter

    def _finish(self, req, out_data, output):
        L = F(self.QUANT_LEVEL)
        self.codes = mx_round(mx_mul(output, L))
        self.assign(out_data[0], req[0], mx_div(self.codes, L))


class QIL_PY(_QILBase):
    def forward(self, is_train, req, in_data, out_data, aux):
        if in_data[1][0] < 0:                                              # QIL.py:51-54 (in place)
            in_data[1][...] = 0.0
        if in_data[2][0] > 1.0:
            in_data[2][...] = 1.0
        assert in_data[1][0] < in_data[2][0], "pruning_point vs clipping_point"   # :56
        x, pp, cp = fl(in_data[0]), fl(in_data[1]), fl(in_data[2])
        center = mx_mul(F(0.5), mx_add(cp, pp))                            # :72
        distance = mx_mul(F(0.5), mx_sub(cp, pp))                          # :73
        a = mx_div(F(0.5), distance)                                       # :74
        b = mx_add(mx_div(mx_mul(F(-0.5), center), distance), F(0.5))      # :75
        output, xabs, sgn, inter = self._transform(x, pp, cp, a, b)        # :76-80
        self._saved = (xabs, sgn, inter, a, center, distance)
        self._finish(req, out_data, output)                                # :85 ("ste" type)

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        assert len(req) >= 3 and self.fix_gamma == True                    # :118-119
        xabs, sgn, inter, a, center, distance = self._saved
        g = mx_mul(mx_mul(fl(out_grad[0]), sgn), inter)                    # d out / d lin
        dx = mx_mul(mx_mul(g, a), sgn)
        da = mx_sum(mx_mul(g, xabs))
        db = mx_sum(g)
        # a = 0.5/d ; b = -0.5 c/d + 0.5 ; c = 0.5(cp+pp) ; d = 0.5(cp-pp)
        dd = mx_add(mx_mul(da, mx_div(F(-0.5), mx_mul(distance, distance))),
                    mx_mul(db, mx_div(mx_mul(F(0.5), center), mx_mul(distance, distance))))
        dc = mx_mul(db, mx_div(F(-0.5), distance))
        dcp = mx_add(mx_mul(F(0.5), dc), mx_mul(F(0.5), dd))
        dpp = mx_sub(mx_mul(F(0.5), dc), mx_mul(F(0.5), dd))
        self.assign(in_grad[0], req[0], dx)
        self.assign(in_grad[1], req[1], dpp.reshape(1))
        self.assign(in_grad[2], req[2], dcp.reshape(1))


class QIL_V2_PY(_QILBase):
    def forward(self, is_train, req, in_data, out_data, aux):
        assert len(in_data) == 4                                           # QIL_V2.py:34
        x, center, distance = fl(in_data[0]), fl(in_data[1]), fl(in_data[2])
        cp = mx_add(center, distance)                                      # :46
        pp = mx_sub(center, distance)                                      # :47
        a = mx_div(F(0.5), distance)                                       # :48
        b = mx_add(mx_div(mx_mul(F(-0.5), center), distance), F(0.5))      # :49
        output, xabs, sgn, inter = self._transform(x, pp, cp, a, b)        # :50-56
        self._saved = (xabs, sgn, inter, a, center, distance)
        self._finish(req, out_data, output)                                # :58-59

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        assert len(req) >= 3 and self.fix_gamma == True
        xabs, sgn, inter, a, center, distance = self._saved
        g = mx_mul(mx_mul(fl(out_grad[0]), sgn), inter)
        dx = mx_mul(mx_mul(g, a), sgn)
        da = mx_sum(mx_mul(g, xabs))
        db = mx_sum(g)
        dd = mx_add(mx_mul(da, mx_div(F(-0.5), mx_mul(distance, distance))),
                    mx_mul(db, mx_div(mx_mul(F(0.5), center), mx_mul(distance, distance))))
        dc = mx_mul(db, mx_div(F(-0.5), distance))
        self.assign(in_grad[0], req[0], dx)
        self.assign(in_grad[1], req[1], dc.reshape(1))
        self.assign(in_grad[2], req[2], dd.reshape(1))


class QIL_V3_PY(_QILBase):
    def forward(self, is_train, req, in_data, out_data, aux):
        x, ep, ed = fl(in_data[0]), fl(in_data[1]), fl(in_data[2])
        pp = np.exp(ep).astype(F)                                          # QIL_V3.py:50
        distance = np.exp(ed).astype(F)                                    # :51
        cp = mx_add(pp, distance)                                          # :52
        xabs = np.abs(x)
        sgn = mx_sign(x)
        inter = ((xabs >= pp).astype(F) * (xabs <= cp).astype(F)).astype(F)  # :56
        lin = mx_div(mx_sub(xabs, pp), distance)                           # :59
        output = mx_add(mx_mul(sgn, (xabs > cp).astype(F)), mx_mul(mx_mul(sgn, lin), inter))  # :58-59
        self._saved = (xabs, sgn, inter, pp, distance)
        self._finish(req, out_data, output)                                # :61

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        assert len(req) >= 3 and self.fix_gamma == True
        xabs, sgn, inter, pp, distance = self._saved
        g = mx_mul(mx_mul(fl(out_grad[0]), sgn), inter)                    # d out / d lin
        dx = mx_mul(mx_div(g, distance), sgn)
        dpp = mx_sum(mx_div(-g, distance))
        ddist = mx_sum(mx_mul(g, mx_div(-mx_sub(xabs, pp), mx_mul(distance, distance))))
        self.assign(in_grad[0], req[0], dx)
        self.assign(in_grad[1], req[1], mx_mul(dpp, pp).reshape(1))        # d exp(ep) = exp(ep)
        self.assign(in_grad[2], req[2], mx_mul(ddist, distance).reshape(1))


# ----------------------------------------------------------------------------------------------
# registry: op_type -> (oracle class, constructor-from-string-attrs)
# ----------------------------------------------------------------------------------------------
def _b(v):
    return v if isinstance(v, bool) else bool(eval(str(v)))


def create(op_type, **attrs):
    """Build an oracle operator from the *string* attributes the reference Props take."""
    a = dict(attrs)
    if op_type in ("Quantization_int8_V2", "ClipGrad_Quantization_int8"):
        cls = Quantization_int8 if op_type == "Quantization_int8_V2" else ClipGrad_Quantization_int8
        return cls(str(a["quant_mode"]), _b(a["is_weight"]), _b(a.get("is_weight_perchannel", "False")),
                   int(a.get("delay_quant", 0)), float(a.get("ema_decay", 0.99)))
    if op_type == "GDRQ_Fold_BN":
        ev = lambda k, d: eval(str(a.get(k, d)))
        return GDRQ_Fold_BN(str(a["quant_mode"]), _b(a.get("is_weight_perchannel", "False")),
                            int(a.get("delay_quant", 0)), float(a.get("ema_decay", 0.99)),
                            str(a.get("name", "fold_bn")), int(a["num_filter"]), int(a["num_group"]),
                            ev("kernel", "(3,3)"), ev("stride", "(1,1)"), ev("pad", "(0,0)"),
                            ev("dilate", "(1,1)"), _b(a.get("no_bias", "True")), float(a.get("eps", 1e-5)),
                            float(a.get("momentum", 0.9)), _b(a.get("fix_gamma", "False")),
                            _b(a.get("quantize_flag", "True")))
    if op_type == "GDRQ_PY":
        return GDRQ_PY(int(a.get("nbits", 4)), int(a.get("group_size", -1)), _b(a.get("is_weight", "False")),
                       float(a.get("lamda", 0.001)), int(a.get("delay_quant", 0)),
                       _b(a.get("fix_alpha", "False")), float(a.get("ktimes", 3)))
    if op_type == "CLIP_RELU_PY":
        return CLIP_RELU_PY(eval(str(a.get("nbits", "8"))), eval(str(a.get("threshold", "8.0"))))
    if op_type == "QUANT_STE_PY":
        return QUANT_STE_PY(eval(str(a.get("nbits", "8"))))
    if op_type == "PACT_PY":
        return PACT_PY(eval(str(a.get("nbits", "8"))))
    if op_type == "PACT_V2_PY":
        return PACT_V2_PY(eval(str(a.get("nbits", "8"))))
    if op_type == "DoReFa_PY":
        return DoReFa_PY(eval(str(a.get("nbits", "8"))))
    if op_type == "WNQ_PY":
        return WNQ_PY(int(a.get("nbits", 4)), _b(a.get("is_perchannel", "False")))
    if op_type in ("QIL_PY", "QIL_V2_PY", "QIL_V3_PY"):
        cls = {"QIL_PY": QIL_PY, "QIL_V2_PY": QIL_V2_PY, "QIL_V3_PY": QIL_V3_PY}[op_type]
        return cls(_b(a.get("is_weight", "False")), _b(a.get("fix_gamma", "True")), int(a.get("nbits", "4")))
    raise KeyError(op_type)
