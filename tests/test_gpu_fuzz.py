"""Randomised CUDA-vs-oracle checks (hypothesis): arbitrary shapes, value scales, views at odd offsets, all req modes,
plus the error behaviour of the boundary (bad arguments raise B2QError / ValueError, nothing is silently copied)."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import quant_oracle as qo
from tests.golden_util import bits_equal

pytestmark = pytest.mark.gpu
F = np.float32
_SET = dict(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck))


def _ops(op_type, **attrs):
    import b200quant
    attrs = {k: str(v) for k, v in attrs.items()}
    return b200quant.get_prop(op_type)(**attrs).create_operator(None, None, None), qo.create(op_type, **attrs)


@given(st.lists(st.integers(1, 9), min_size=1, max_size=4), st.sampled_from([1e-3, 0.3, 1.0, 40.0]),
       st.sampled_from(["Quantization_int8_V2", "ClipGrad_Quantization_int8"]), st.booleans(), st.booleans(),
       st.sampled_from(["write", "add"]), st.integers(0, 7), st.integers(0, 2 ** 31 - 1))
@settings(**_SET)
def test_minmax_random(shape, scale, op_type, is_weight, per_channel, req, offset, seed):
    import torch
    rng = np.random.default_rng(seed)
    shape = tuple(shape)
    per_channel = per_channel and is_weight
    op, ref = _ops(op_type, quant_mode="minmax", is_weight=is_weight, is_weight_perchannel=per_channel)
    n = int(np.prod(shape))
    naux = shape[0] if per_channel else 1
    x = (rng.standard_normal(shape) * scale).astype(F)
    x.flat[0] = F(scale)                                        # never all zero
    y0 = rng.standard_normal(shape).astype(F)
    # device tensors are views at an odd element offset inside larger buffers (exercises the unaligned paths)
    xbuf = torch.zeros(n + 16, device="cuda")
    ybuf = torch.zeros(n + 16, device="cuda")
    xd = xbuf[offset:offset + n].view(shape)
    yd = ybuf[offset:offset + n].view(shape)
    xd.copy_(torch.from_numpy(x))
    yd.copy_(torch.from_numpy(y0))
    aux_d, aux_r = [torch.full((naux,), 0.9, device="cuda")], [np.full(naux, 0.9, F)]
    yr = y0.copy()
    for train in (True, False):
        op.forward(train, [req], [xd], [yd], aux_d)
        ref.forward(train, [req], [x], [yr], aux_r)
        assert bits_equal(aux_d[0].cpu().numpy(), aux_r[0])
        assert bits_equal(yd.cpu().numpy(), yr)
    assert float(ybuf[:offset].abs().sum()) == 0 and float(ybuf[offset + n:].abs().sum()) == 0   # no out-of-range writes


@given(st.integers(1, 4), st.integers(1, 6), st.integers(1, 3), st.integers(1, 7), st.integers(1, 7), st.booleans(),
       st.integers(0, 2 ** 31 - 1))
@settings(**_SET)
def test_gdrq_grouped_random(n, groups, gs, h, w, is_weight, seed):
    import torch
    rng = np.random.default_rng(seed)
    c = groups * gs
    shape = (c, n, h, w) if is_weight else (n, c, h, w)
    op, ref = _ops("GDRQ_PY", nbits=6, group_size=gs, is_weight=is_weight, lamda=0.01, delay_quant=0, fix_alpha=False,
                   ktimes=2)
    x = rng.standard_normal(shape).astype(F)
    dy = rng.standard_normal(shape).astype(F)
    aux_d, aux_r = [torch.ones(groups, device="cuda")], [np.ones(groups, F)]
    xd, yd, yr = torch.from_numpy(x).cuda(), torch.zeros(shape, device="cuda"), np.zeros(shape, F)
    op.forward(True, ["write"], [xd], [yd], aux_d)
    ref.forward(True, ["write"], [x], [yr], aux_r)
    np.testing.assert_allclose(aux_d[0].cpu().numpy(), aux_r[0], rtol=1e-6)
    if bits_equal(aux_d[0].cpu().numpy(), aux_r[0]):
        assert bits_equal(yd.cpu().numpy(), yr)
    gd, gr = torch.zeros(shape, device="cuda"), np.zeros(shape, F)
    op.backward(["write"], [torch.from_numpy(dy).cuda()], [xd], [yd], [gd], aux_d)
    ref.backward(["write"], [dy], [x], [yr], [gr], [aux_d[0].cpu().numpy()])
    assert bits_equal(gd.cpu().numpy(), gr)


def test_bad_arguments_raise():
    import torch
    import b200quant
    from b200quant import _kernels as K
    from b200quant._lib import B2QError
    op = b200quant.get_prop("Quantization_int8_V2")(quant_mode="minmax", is_weight="False").create_operator(None, None, None)
    x = torch.randn(4, 4, device="cuda")
    with pytest.raises(ValueError):      # wrong aux size
        op.forward(True, ["write"], [x], [torch.empty_like(x)], [torch.ones(3, device="cuda")])
    with pytest.raises(ValueError):      # output size mismatch
        op.forward(True, ["write"], [x], [torch.empty(5, device="cuda")], [torch.ones(1, device="cuda")])
    with pytest.raises(TypeError):       # not float32
        op.forward(True, ["write"], [x.double()], [torch.empty_like(x)], [torch.ones(1, device="cuda")])
    with pytest.raises(ValueError):      # non-contiguous view: never copied silently
        op.forward(True, ["write"], [x.t()], [torch.empty_like(x)], [torch.ones(1, device="cuda")])
    with pytest.raises(ValueError):      # unknown req
        op.forward(True, ["overwrite"], [x], [torch.empty_like(x)], [torch.ones(1, device="cuda")])
    with pytest.raises(ValueError):      # mixed host / device tensors
        op.forward(True, ["write"], [x], [torch.empty(4, 4)], [torch.ones(1, device="cuda")])
    with pytest.raises(B2QError):        # C ABI argument check surfaces with the library's message
        K.qdq(x, torch.empty_like(x), torch.ones(1, device="cuda"), 127, 99, "write")
    op.forward(True, ["null"], [x], [torch.empty_like(x)], [torch.ones(1, device="cuda")])   # null req is legal


@pytest.mark.parametrize("op_type", ["Quantization_int8_V2", "ClipGrad_Quantization_int8", "GDRQ_PY"])
def test_guard_regions_around_medium_tensors(op_type):
    """Out-of-range WRITES of the vectorised paths: tensors of 2^16 + k elements placed at every float offset 0..8 inside
    buffers whose surroundings hold a canary; forward and backward must leave every canary untouched and match the
    oracle bit for bit (compute-sanitizer is not available on the GPU pool, this is the bounds check we run instead)."""
    import torch
    rng = np.random.default_rng(31)
    canary = 12345.678
    pad = 64
    for k in (0, 1, 7, 8, 9, 31, 255):
        n = (1 << 16) + k
        x = (rng.standard_normal(n) * 2).astype(F)
        dy = rng.standard_normal(n).astype(F)
        for offset in (0, 1, 3, 8):
            if op_type == "GDRQ_PY":
                op, ref = _ops(op_type, nbits=8, group_size=-1, is_weight=False, lamda=0.01, ktimes=3)
            else:
                op, ref = _ops(op_type, quant_mode="minmax", is_weight=False)
            bufs = [torch.full((n + 2 * pad + 16,), canary, device="cuda") for _ in range(4)]
            lo = pad + offset
            xd, yd, gd, dd = (b[lo:lo + n] for b in bufs)
            if k == 9:   # output misaligned against the input: the scalar fallback paths
                bufs[1] = torch.full((n + 2 * pad + 16,), canary, device="cuda")
                yd = bufs[1][lo + 1:lo + 1 + n]
            xd.copy_(torch.from_numpy(x))
            gd.copy_(torch.from_numpy(dy))
            aux_d, aux_r = [torch.ones(1, device="cuda")], [np.ones(1, F)]
            yr, dr = np.zeros(n, F), np.zeros(n, F)
            op.forward(True, ["write"], [xd], [yd], aux_d)
            ref.forward(True, ["write"], [x], [yr], aux_r)
            op.backward(["write"], [gd], [xd], [yd], [dd], aux_d)
            ref.backward(["write"], [dy], [x], [yr], [dr], aux_r)
            assert bits_equal(aux_d[0].cpu().numpy(), aux_r[0]), (k, offset)
            assert bits_equal(yd.cpu().numpy(), yr), (k, offset)
            assert bits_equal(dd.cpu().numpy(), dr), (k, offset)
            for i, b in enumerate(bufs):
                a = lo + 1 if (k == 9 and i == 1) else lo
                assert bool((b[:a] == canary).all()) and bool((b[a + n:] == canary).all()), (k, offset, i)


@pytest.mark.parametrize("shape,gs", [((64, 64, 14, 14), 1), ((64, 64, 14, 14), 4), ((128, 96, 7, 7), 1), ((128, 96, 7, 7), 8),
                                      ((96, 40, 9, 11), 5), ((40, 48, 13, 13), 2), ((33, 40, 9, 11), 1)])
@pytest.mark.parametrize("req", ["write", "add"])
def test_gdrq_grouped_activations_on_small_feature_maps(shape, gs, req):
    """Grouped GDRQ_PY activations (GDRQ.py:88-118, :131-152) whose rows are not a multiple of eight floats -- 14x14, 7x7 and
    odd maps: large tensors take the flattened (row, element) kernels (reduce_seg_kernel<.., 2|3>, qdq_seg_flatidx_kernel,
    bwd_seg_flatidx_kernel); the last shape stays below their size limit and takes the row-at-a-time kernels.  Two
    training steps against the oracle: alpha to 1e-6, and bit-exact output / gradient given the same alpha."""
    import torch
    rng = np.random.default_rng(hash((shape, gs)) % (1 << 31))
    groups = shape[1] // gs
    op, ref = _ops("GDRQ_PY", nbits=8, group_size=gs, is_weight=False, lamda=0.001, delay_quant=0, fix_alpha=False, ktimes=3)
    aux_d, aux_r = [torch.ones(groups, device="cuda")], [np.ones(groups, F)]
    for step in range(2):
        x = (rng.standard_normal(shape) * (0.5 + step)).astype(F)
        dy = rng.standard_normal(shape).astype(F)
        xd, yd, yr = torch.from_numpy(x).cuda(), torch.zeros(shape, device="cuda"), np.zeros(shape, F)
        op.forward(True, ["write"], [xd], [yd], aux_d)
        ref.forward(True, ["write"], [x], [yr], aux_r)
        np.testing.assert_allclose(aux_d[0].cpu().numpy(), aux_r[0], rtol=1e-6)
        aux_r[0][...] = aux_d[0].cpu().numpy()        # same alpha on both sides from here on
        c = np.where(np.abs(x) <= aux_r[0][None, :, None, None].repeat(gs, 1), x,
                     (aux_r[0][None, :, None, None].repeat(gs, 1) * np.sign(x)).astype(F)).astype(F)
        q = qo.mx_div(np.broadcast_to(aux_r[0][None, :, None, None].repeat(gs, 1), shape).astype(F), F(255))
        want, _ = qo.qdq(c, q)
        assert bits_equal(yd.cpu().numpy(), want)
        g0 = rng.standard_normal(shape).astype(F)
        gd, gr = torch.from_numpy(g0.copy()).cuda(), g0.copy()
        op.backward([req], [torch.from_numpy(dy).cuda()], [xd], [yd], [gd], aux_d)
        ref.backward([req], [dy], [x], [want], [gr], [aux_r[0]])
        assert bits_equal(gd.cpu().numpy(), gr)


@pytest.mark.parametrize("shape", [(64, 64, 14, 14), (128, 96, 7, 7), (96, 40, 9, 11), (40, 48, 13, 13), (33, 40, 9, 11),
                                   (70, 64, 12, 12), (64, 32, 28, 28)])
def test_grouped_statistics_on_small_feature_maps(shape):
    """b2q_absmax_f32 / b2q_meanabs_f32 per channel group on an (N, C, H, W) view whose rows are short and mostly not a
    multiple of eight floats (flattened-walk kernels for the large ones): max|x| exactly, mean|x| = the correctly rounded
    float32 of the exact sum divided by float32(count), as the oracle's mx_mean."""
    import torch
    from b200quant import _kernels as K
    rng = np.random.default_rng(sum(shape))
    x = (rng.standard_normal(shape) * 1.7).astype(F)
    xd = torch.from_numpy(x).cuda()
    for gs in (1, 4, 8):
        if shape[1] % gs:
            continue
        groups = shape[1] // gs
        view = (shape[0], groups, gs * shape[2] * shape[3])
        mx, mean = torch.zeros(groups, device="cuda"), torch.zeros(groups, device="cuda")
        K.absmax(xd, mx, view)
        K.meanabs(xd, mean, view)
        xg = np.abs(x).reshape(shape[0], groups, -1)
        assert bits_equal(mx.cpu().numpy(), xg.max(axis=(0, 2)))
        want = np.array([qo.mx_mean(xg[:, g, :]) for g in range(groups)], F)
        assert bits_equal(mean.cpu().numpy(), want)
