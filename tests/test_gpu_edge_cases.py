"""Adversarial inputs of SURVEY.md section 8d -- all-zero tensors, NaN, +-Inf, denormals, one huge outlier -- through
every first-tier operator, on the whole-tensor (flat) kernels, the segmented (grouped / per-channel) kernels and the
row-fused weight kernels, against the NumPy oracle.

What the reference does with them (and what is asserted here, element by element):
  * all-zero weight          max|w| = 0 -> q = 0 -> 0/0 = NaN everywhere (quant_ops.py:26-28)
  * NaN element              mx.nd.clip passes it through, max|x| becomes NaN, so the threshold and with it the whole
                             output turn NaN: a diverged network is not silently quantised
  * +-Inf element            max|x| = Inf -> q = Inf -> finite/Inf = 0 -> 0*Inf = NaN, Inf/Inf = NaN
  * denormals                no flush to zero anywhere (the library is built without -use_fast_math)
  * one huge outlier         every other element rounds to a signed zero

NaN payloads differ between x86 (0xffc00000) and the GPU (0x7fffffff); the comparison requires "NaN at the same
positions" and bit equality everywhere else (signed zeros included)."""
import numpy as np
import pytest

from oracle import quant_oracle as qo

pytestmark = pytest.mark.gpu
F = np.float32


@pytest.fixture(scope="module")
def T():
    import torch
    return torch


def dev(T, a):
    return T.from_numpy(np.ascontiguousarray(a, dtype=F)).cuda()


def host(t):
    return t.detach().cpu().numpy()


def make(op_type, **attrs):
    import b200quant
    attrs = {k: str(v) for k, v in attrs.items()}
    return b200quant.get_prop(op_type)(**attrs).create_operator(None, None, None), qo.create(op_type, **attrs)


def same(a, b):
    a = np.ascontiguousarray(a, dtype=F)
    b = np.ascontiguousarray(b, dtype=F)
    if a.shape != b.shape:
        return False
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return False
    return np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb])


def poison(kind, shape, rng):
    """float32 tensor of ``shape`` carrying one adversarial feature (position chosen away from element 0 on purpose:
    the per-block / per-warp reductions must find it wherever it is)."""
    n = int(np.prod(shape))
    x = rng.standard_normal(n).astype(F)
    at = (n * 5) // 7
    if kind == "zeros":
        x[:] = 0
    elif kind == "neg_zeros":
        x[:] = -0.0
    elif kind == "nan":
        x[at] = np.nan
    elif kind == "inf":
        x[at] = np.inf
    elif kind == "neg_inf":
        x[at] = -np.inf
    elif kind == "denormal":
        x = (x * F(1e-41)).astype(F)           # everything subnormal
    elif kind == "mixed_denormal":
        x[::3] = (x[::3] * F(1e-42)).astype(F)
    elif kind == "outlier":
        x[at] = F(3e30)
    elif kind == "tiny_outlier":
        x = (x * F(1e-30)).astype(F)
        x[at] = F(1.0)
    else:
        raise ValueError(kind)
    return x.reshape(shape)


KINDS = ["zeros", "neg_zeros", "nan", "inf", "neg_inf", "denormal", "mixed_denormal", "outlier", "tiny_outlier"]
ACT_SHAPES = [(4, 16, 14, 14), (4099,), (2, 3, 64, 64)]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("op_type", ["Quantization_int8_V2", "ClipGrad_Quantization_int8"])
def test_minmax_activation_edge_inputs(T, op_type, kind):
    rng = np.random.default_rng(31)
    for shape in ACT_SHAPES:
        op, ref = make(op_type, quant_mode="minmax", is_weight=False, is_weight_perchannel=False, delay_quant=0,
                       ema_decay=0.99)
        aux_d, aux_r = [dev(T, np.ones(1, F))], [np.ones(1, F)]
        # a clean batch, the poisoned one, a clean one again (the state machine must carry what the reference carries)
        for step, k in enumerate([None, kind, None]):
            x = poison(k, shape, rng) if k else rng.standard_normal(shape).astype(F)
            dy = rng.standard_normal(shape).astype(F)
            xd, yd, yr = dev(T, x), dev(T, np.zeros(shape, F)), np.zeros(shape, F)
            op.forward(True, ["write"], [xd], [yd], aux_d)
            with np.errstate(all="ignore"):
                ref.forward(True, ["write"], [x], [yr], aux_r)
            assert same(host(aux_d[0]), aux_r[0]), (shape, step, "aux", host(aux_d[0]), aux_r[0])
            assert same(host(yd), yr), (shape, step, "out")
            gd, gr = dev(T, np.zeros(shape, F)), np.zeros(shape, F)
            op.backward(["write"], [dev(T, dy)], [xd], [yd], [gd], aux_d)
            with np.errstate(all="ignore"):
                ref.backward(["write"], [dy], [x], [yr], [gr], aux_r)
            assert same(host(gd), gr), (shape, step, "grad")


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("per_channel", [False, True])
@pytest.mark.parametrize("op_type", ["Quantization_int8_V2", "ClipGrad_Quantization_int8"])
def test_minmax_weight_edge_inputs(T, op_type, per_channel, kind):
    rng = np.random.default_rng(32)
    # short rows (row-fused warp kernel), long rows (CTA-per-row / segmented), rows of 9 (depthwise)
    for shape in [(16, 8, 3, 3), (4, 2048, 3, 3), (33, 1, 3, 3)]:
        op, ref = make(op_type, quant_mode="minmax", is_weight=True, is_weight_perchannel=per_channel, delay_quant=0,
                       ema_decay=0.99)
        naux = shape[0] if per_channel else 1
        aux_d, aux_r = [dev(T, np.ones(naux, F))], [np.ones(naux, F)]
        x = poison(kind, shape, rng)
        yd, yr = dev(T, np.zeros(shape, F)), np.zeros(shape, F)
        op.forward(True, ["write"], [dev(T, x)], [yd], aux_d)
        with np.errstate(all="ignore"):
            ref.forward(True, ["write"], [x], [yr], aux_r)
        assert same(host(aux_d[0]), aux_r[0]), (shape, "aux")
        assert same(host(yd), yr), (shape, "out")


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("group_size", [-1, 4])
@pytest.mark.parametrize("is_weight", [False, True])
def test_gdrq_edge_inputs(T, is_weight, group_size, kind):
    rng = np.random.default_rng(33)
    shape = (16, 8, 3, 3) if is_weight else (4, 16, 14, 14)
    op, ref = make("GDRQ_PY", nbits=8, group_size=group_size, is_weight=is_weight, lamda=0.001, delay_quant=0,
                   fix_alpha=False, ktimes=3)
    ch = shape[0] if is_weight else shape[1]
    g = 1 if group_size == -1 else ch // group_size
    aux_d, aux_r = [dev(T, np.ones(g, F))], [np.ones(g, F)]
    x = poison(kind, shape, rng)
    dy = rng.standard_normal(shape).astype(F)
    xd, yd, yr = dev(T, x), dev(T, np.zeros(shape, F)), np.zeros(shape, F)
    op.forward(True, ["write"], [xd], [yd], aux_d)
    with np.errstate(all="ignore"):
        ref.forward(True, ["write"], [x], [yr], aux_r)
    a_d, a_r = host(aux_d[0]), aux_r[0]
    assert np.array_equal(np.isnan(a_d), np.isnan(a_r))
    ok = ~np.isnan(a_r)
    np.testing.assert_allclose(a_d[ok], a_r[ok], rtol=1e-6)
    if same(a_d, a_r):
        assert same(host(yd), yr)
        gd, gr = dev(T, np.zeros(shape, F)), np.zeros(shape, F)
        op.backward(["write"], [dev(T, dy)], [xd], [yd], [gd], aux_d)
        with np.errstate(all="ignore"):
            ref.backward(["write"], [dy], [x], [yr], [gr], aux_r)
        assert same(host(gd), gr)


@pytest.mark.parametrize("kind", KINDS)
def test_foldbn_data_and_weight_edge_inputs(T, kind):
    """GDRQ_Fold_BN's two quantisation paths (fold_bn_v1_gdrq.py:53-96) via the array-level calls the operator makes."""
    import b200quant._kernels as K
    rng = np.random.default_rng(34)
    x = poison(kind, (4, 8, 14, 14), rng)
    ref = qo.create("GDRQ_Fold_BN", quant_mode="minmax", is_weight_perchannel="True", delay_quant="0", ema_decay="0.99",
                    name="f", num_filter="16", num_group="1", kernel="(3, 3)", stride="(1, 1)", pad="(1, 1)",
                    dilate="(1, 1)", no_bias="True", eps="1e-5", momentum="0.9", fix_gamma="False", quantize_flag="True")
    w = poison(kind, (16, 8, 3, 3), rng)
    gamma, beta = rng.uniform(0.5, 1.5, 16).astype(F), rng.standard_normal(16).astype(F)
    mean, var = rng.standard_normal(16).astype(F), rng.uniform(0.5, 1.5, 16).astype(F)
    bn_out = rng.standard_normal((4, 16, 14, 14)).astype(F)
    aux_r = [np.ones(1, F), np.ones(16, F)]
    yr = np.zeros((4, 16, 14, 14), F)
    with np.errstate(all="ignore"):
        ref.forward(True, ["write"], [x, w, bn_out, gamma, beta, mean, var], [yr], aux_r)
    xd, xq, a0 = dev(T, x), dev(T, np.zeros_like(x)), dev(T, np.ones(1, F))
    K.foldbn_data_fwd(xd, xq, a0, True, 0.99)
    assert same(host(a0), aux_r[0]) and same(host(xq), ref.data_q)
    wq, bias, a1 = dev(T, np.zeros_like(w)), dev(T, np.zeros(16, F)), dev(T, np.ones(16, F))
    K.foldbn_weight_fwd(dev(T, w), wq, bias, a1, dev(T, gamma), dev(T, beta), dev(T, mean), dev(T, var), 1e-5, True,
                        True, True)
    assert same(host(a1), aux_r[1]) and same(host(wq), ref.weight_q) and same(host(bias), ref.bias)


def test_nan_is_not_swallowed_by_any_clip_mode(T):
    """fminf/fmaxf drop NaN, so the fast clip path used to turn a NaN input into -T (0 for the ReLU clip)
    before the safety test ran.  Every clip mode, flat and segmented kernels, NaN in a word of otherwise safe values."""
    import b200quant._kernels as K
    from b200quant import _lib
    rng = np.random.default_rng(35)
    n = 8 * 1024 + 5
    x = rng.standard_normal(n).astype(F)
    x[[3, 1000, 4097, n - 1]] = np.nan
    thr = np.array([1.25], F)
    for clip in (_lib.CLIP_NONE, _lib.CLIP_SYM, _lib.CLIP_WHERE_LE, _lib.CLIP_ZERO_T, _lib.CLIP_PACT, _lib.CLIP_WHERE_LT):
        with np.errstate(all="ignore"):
            c = qo.clip_by_mode(clip, x, thr[0])
            want, _ = qo.qdq(c, qo.mx_div(thr, F(127)))
        yd = dev(T, np.zeros(n, F))
        K.qdq(dev(T, x), yd, dev(T, thr), 127, clip, "write")
        assert same(host(yd), want), clip
    # segmented hot kernel (grouped activation, > 2^20 elements so the 256-bit per-piece loop runs)
    xs = rng.standard_normal((8, 16, 96, 96)).astype(F)
    xs[3, 5, 7, 9] = np.nan
    thr4 = np.array([1.0, 1.5, 0.5, 2.0], F)
    for clip in (_lib.CLIP_SYM, _lib.CLIP_WHERE_LE):
        with np.errstate(all="ignore"):
            r = xs.reshape(8, 4, -1)
            c = np.stack([qo.clip_by_mode(clip, r[:, g], thr4[g]) for g in range(4)], axis=1)
            want = np.stack([qo.qdq(c[:, g], qo.mx_div(thr4[g], F(127)))[0] for g in range(4)], axis=1).reshape(xs.shape)
        yd = dev(T, np.zeros(xs.shape, F))
        K.qdq(dev(T, xs), yd, dev(T, thr4), 127, clip, "write", view=(8, 4, 4 * 96 * 96))
        assert same(host(yd), want), clip
