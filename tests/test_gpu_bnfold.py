"""batch statistics of the convolution output (BatchNorm_v1 with output_mean_var) and the fold-BN
weight path that consumes them, in one launch -- against the oracle.  Two-stage parity: mean / var
are mean-derived quantities (1e-6 relative, BASELINE.json north_star); given the SAME mean / var the folded, quantised
weight, its per-channel thresholds and the bias are bit-exact."""
import numpy as np
import pytest

from oracle import quant_oracle as qo
from tests.golden_util import bits_equal

pytestmark = pytest.mark.gpu
F = np.float32


def dev(T, a):
    return T.from_numpy(np.ascontiguousarray(a, dtype=F)).cuda()


@pytest.mark.parametrize("shape", [(8, 16, 14, 14), (4, 32, 7, 7), (3, 5, 9, 11), (256, 64, 56, 56), (32, 1024, 7, 7),
                                   (2, 8, 1, 1)])
def test_batch_stats_match_batchnorm_v1(shape):
    import torch as T
    import b200quant._kernels as K
    g = T.Generator(device="cuda").manual_seed(3)
    y = T.empty(shape, device="cuda").normal_(0.3, 1.7, generator=g)
    mean, var = T.empty(shape[1], device="cuda"), T.empty(shape[1], device="cuda")
    K.bn_batch_stats(y, mean, var)
    m_r, v_r = qo.bn_v1_batch_stats(y.cpu().numpy())
    np.testing.assert_allclose(mean.cpu().numpy(), m_r, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(var.cpu().numpy(), v_r, rtol=1e-6)
    # deterministic: fixed partition, fixed combination order
    m2, v2 = T.empty_like(mean), T.empty_like(var)
    K.bn_batch_stats(y, m2, v2)
    assert T.equal(mean, m2) and T.equal(var, v2)
    # and it is the library's biased batch variance
    want = y.double().var(dim=(0, 2, 3), unbiased=False)
    T.testing.assert_close(var.double(), want, rtol=2e-6, atol=0)


@pytest.mark.parametrize("cfg", [dict(n=8, cin=16, cout=32, k=3, group=1, hw=14, pc=True),
                                 dict(n=8, cin=32, cout=32, k=3, group=32, hw=14, pc=True),       # depthwise: rows of 9
                                 dict(n=4, cin=64, cout=128, k=1, group=1, hw=7, pc=True),
                                 dict(n=4, cin=16, cout=24, k=3, group=1, hw=9, pc=False),        # per-tensor: two launches
                                 dict(n=2, cin=512, cout=512, k=3, group=1, hw=4, pc=True)])      # rows of 4608
@pytest.mark.parametrize("is_train", [True, False])
def test_fused_batchstat_fold_quantise(cfg, is_train):
    import torch as T
    import b200quant._kernels as K
    from b200quant import _lib
    rng = np.random.default_rng(11)
    n, cin, cout, k, hw = cfg["n"], cfg["cin"], cfg["cout"], cfg["k"], cfg["hw"]
    w = (rng.standard_normal((cout, cin // cfg["group"], k, k)) * 0.2).astype(F)
    conv_out = (rng.standard_normal((n, cout, hw, hw)) * 1.3 + 0.2).astype(F)
    gamma, beta = rng.uniform(0.5, 1.5, cout).astype(F), rng.standard_normal(cout).astype(F)
    naux = cout if cfg["pc"] else 1
    wq, bias, aux = dev(T, np.zeros_like(w)), dev(T, np.zeros(cout, F)), dev(T, np.full(naux, 7.0, F))
    mean, var = dev(T, np.zeros(cout, F)), dev(T, np.zeros(cout, F))
    ctx = _lib.context(0)
    l0 = ctx.launch_count()
    K.bnstat_foldbn_weight_fwd(dev(T, conv_out), mean, var, dev(T, w), wq, bias, aux, dev(T, gamma), dev(T, beta), 1e-5,
                               cfg["pc"], True, is_train)
    launches = ctx.launch_count() - l0
    assert launches == (1 if cfg["pc"] else 3)            # per-channel: ONE launch for statistics + fold + QDQ + bias
    m_r, v_r = qo.bn_v1_batch_stats(conv_out)
    np.testing.assert_allclose(mean.cpu().numpy(), m_r, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(var.cpu().numpy(), v_r, rtol=1e-6)
    # stage 2: the oracle's fold-BN weight path (fold_bn_v1_gdrq.py:70-96,113; oracle/c is bit-equal to the NumPy oracle,
    # tests/test_oracle_c.py) fed the statistics the kernel produced -> bit-exact
    from oracle import c_oracle as co
    wq_r, bias_r, aux_r = np.zeros_like(w), np.zeros(cout, F), np.full(naux, 7.0, F)
    co.foldbn_weight_fwd(w, wq_r, bias_r, aux_r, gamma, beta, mean.cpu().numpy(), var.cpu().numpy(), 1e-5, cfg["pc"], True,
                         is_train)
    assert bits_equal(wq.cpu().numpy(), wq_r)
    assert bits_equal(bias.cpu().numpy(), bias_r)
    assert bits_equal(aux.cpu().numpy(), aux_r)
    if not is_train:
        assert float(aux.min()) == 7.0        # thresholds are stored only when training (:94-95)
    # and it equals the stand-alone weight entry point on the same statistics
    wq2, bias2, aux2 = dev(T, np.zeros_like(w)), dev(T, np.zeros(cout, F)), dev(T, np.full(naux, 7.0, F))
    K.foldbn_weight_fwd(dev(T, w), wq2, bias2, aux2, dev(T, gamma), dev(T, beta), mean, var, 1e-5, cfg["pc"], True, is_train)
    assert T.equal(wq.view(T.int32), wq2.view(T.int32)) and T.equal(bias.view(T.int32), bias2.view(T.int32))
    assert T.equal(aux.view(T.int32), aux2.view(T.int32))


@pytest.mark.parametrize("group", [1, 16])
def test_foldbn_layer_twin_fused_equals_unfused_and_trains(group):
    """harness.FoldBNConv2d (GDRQ_fold_bn, fold_bn_v1_gdrq.py:237-288): the fused statistics+fold launch and the
    operator's own weight path give the same bits; the gradient reaches weight / gamma / beta through bn_output."""
    import torch as T
    from b200quant.harness import FoldBNConv2d, export_mx_params
    outs = []
    for fused in (True, False):
        T.manual_seed(5)
        layer = FoldBNConv2d("stage1", 16, 16, (3, 3), pad=(1, 1), num_group=group, is_weight_perchannel=True,
                             fused=fused).cuda().train()
        x = (T.rand(4, 16, 10, 10, device="cuda", generator=T.Generator(device="cuda").manual_seed(1)) * 2 - 1)
        x.requires_grad_(True)
        ys = []
        for _ in range(2):            # first batch initialises data_minmax, second takes the EMA branch
            y = layer(x)
            ys.append(y.detach().clone())
        y.square().mean().backward()
        outs.append((ys, layer.weight.grad.clone(), layer.gamma.grad.clone(), layer.beta.grad.clone(), x.grad.clone(),
                     export_mx_params(layer)[1]))
        assert T.isfinite(layer.weight.grad).all() and float(layer.weight.grad.abs().sum()) > 0
        assert float(layer.moving_var.mean()) != 1.0
    (ya, wa, ga, ba, xa, auxa), (yb, wb, gb, bb, xb, auxb) = outs
    for a, b in zip(ya, yb):
        assert T.equal(a.view(T.int32), b.view(T.int32))
    for a, b in ((wa, wb), (ga, gb), (ba, bb), (xa, xb)):     # library convolution backward: not bitwise reproducible
        T.testing.assert_close(a, b, rtol=1e-4, atol=1e-6)
    assert sorted(auxa) == sorted(auxb) == ["stage1_batchnorm_moving_mean", "stage1_batchnorm_moving_var",
                                            "stage1_fold_bn_data_minmax", "stage1_fold_bn_weight_minmax"]
    for k in auxa:
        assert T.equal(auxa[k], auxb[k]), k


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 6])
def test_batch_stats_conversion_paths_agree_on_special_values(variant):
    """The statistics kernel converts float32 -> float64 on the integer pipe (option stream_icvt, default on; exact for
    every finite value, b2q_selftest(5)) and falls back to the conversion instruction for words holding Inf / NaN.
    Channels of zeros, negative zeros, denormals, huge values, Inf and NaN: both paths give the same bits, for every
    tuning variant of the kernel, and equal the double-precision NumPy statistics where those are finite."""
    import torch as T
    import b200quant._kernels as K
    from b200quant import _lib
    rng = np.random.default_rng(17)
    y = rng.standard_normal((6, 12, 16, 16)).astype(F)
    y[:, 1] = 0.0
    y[:, 2] = -0.0
    y[:, 3] = (rng.integers(1, 1 << 22, size=(6, 16, 16)).astype(np.uint32)).view(F)            # denormals
    y[:, 4] = y[:, 4] * F(1e30)
    y[2, 5, 3, 3] = np.inf
    y[1, 6, 0, 0] = np.nan
    y[:, 7, ::2] = 0.0                                                                           # half zeros (post-ReLU)
    y[0, 8, 0, 0] = -np.inf
    y[:, 9] = F(3.0)                                                                             # constant: variance 0
    ctx = _lib.context(0)
    saved = {k: ctx.get_option(k) for k in ("stream_icvt", "bn_variant")}
    got = []
    try:
        ctx.set_option("bn_variant", variant)
        for icvt in (1, 2, 0):
            ctx.set_option("stream_icvt", icvt)
            mean, var = T.empty(12, device="cuda"), T.empty(12, device="cuda")
            K.bn_batch_stats(dev(T, y), mean, var)
            got.append((mean.cpu().numpy(), var.cpu().numpy()))
    finally:
        for k, v in saved.items():
            ctx.set_option(k, v)
    for other in got[1:]:
        assert bits_equal(got[0][0], other[0]) and bits_equal(got[0][1], other[1])
    mean, var = got[0]
    yd = y.astype(np.float64)
    finite = [0, 1, 2, 3, 7, 9, 10, 11]
    want_m = yd.mean(axis=(0, 2, 3))
    want_v = yd.var(axis=(0, 2, 3))
    np.testing.assert_allclose(mean[finite], want_m[finite], rtol=1e-6, atol=1e-37)
    np.testing.assert_allclose(var[finite], want_v[finite], rtol=2e-6, atol=1e-37)
    assert mean[1] == 0.0 and var[1] == 0.0 and mean[2] == 0.0 and var[9] == 0.0 and mean[9] == 3.0
    assert np.isinf(mean[5]) and np.isnan(mean[6]) and np.isnan(var[6]) and mean[8] == -np.inf


def test_tma_staged_ring_gives_the_same_bits_as_the_register_loops():
    """option stream_reduce=1 routes the batch statistics and the grouped mean|x| / max|x| reductions through a ring of
    16 KB shared-memory stages filled by bulk asynchronous copies (off by default: measured slower).  Same bits as the
    default path: rows longer than a stage, rows packed several to a stage, a ragged last stage, 2..6 stages."""
    import torch as T
    import b200quant._kernels as K
    from b200quant import _lib
    ctx = _lib.context(0)
    saved = {k: ctx.get_option(k) for k in ("stream_reduce", "stream_stages", "stream_icvt")}
    g = T.Generator(device="cuda").manual_seed(9)
    cases = [(8, 16, 96, 96), (32, 24, 16, 16), (5, 7, 40, 40), (64, 8, 56, 56)]
    try:
        for shape in cases:
            y = T.empty(shape, device="cuda").normal_(0.2, 1.5, generator=g)
            c = shape[1]
            view = (shape[0], c, shape[2] * shape[3])
            ref = None
            for cfg in ((0, 4, 0), (1, 2, 0), (1, 3, 1), (1, 4, 0), (1, 6, 1)):
                ctx.set_option("stream_reduce", cfg[0])
                ctx.set_option("stream_stages", cfg[1])
                ctx.set_option("stream_icvt", cfg[2])
                mean, var = T.empty(c, device="cuda"), T.empty(c, device="cuda")
                K.bn_batch_stats(y, mean, var)
                ma, mx = T.zeros(c, device="cuda"), T.zeros(c, device="cuda")
                K.meanabs(y, ma, view)
                K.absmax(y, mx, view)
                got = [t.clone() for t in (mean, var, ma, mx)]
                if ref is None:
                    ref = got
                    assert T.equal(mx, y.abs().amax(dim=(0, 2, 3)))
                else:
                    for a, b in zip(ref[2:], got[2:]):                 # mean|x| (correctly rounded sum) and max|x|: same bits
                        assert T.equal(a.view(T.int32), b.view(T.int32)), (shape, cfg)
                    T.testing.assert_close(got[0], ref[0], rtol=1e-6, atol=1e-7)      # a different summation order:
                    T.testing.assert_close(got[1], ref[1], rtol=1e-6, atol=0)         # mean-derived tolerance
    finally:
        for k, v in saved.items():
            ctx.set_option(k, v)
