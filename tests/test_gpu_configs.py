"""BASELINE.json configs 2-5 as parity cases: every quantization node of the reference networks (shapes from
b200quant.workloads, reduced batch so the NumPy oracle finishes in seconds) through the CUDA operators vs the oracle.

config 2  ResNet-50, Quantization_int8_V2 per-tensor weight + activation
config 3  MobileNet-v1 fold-BN (GDRQ_Fold_BN), per-channel weights incl. the 13 depthwise layers (rows of 9)
config 4  MobileNet-v1 GDRQ_PY on weights and activations
config 5  ResNeXt-101 32x4d, ClipGrad_Quantization_int8 forward + masked backward
"""
import numpy as np
import pytest

from oracle import quant_oracle as qo
from tests.golden_util import bits_equal

pytestmark = pytest.mark.gpu
F = np.float32


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dtype=F)).cuda()


def _pair(op_type, **attrs):
    import b200quant
    attrs = {k: str(v) for k, v in attrs.items()}
    return b200quant.get_prop(op_type)(**attrs).create_operator(None, None, None), qo.create(op_type, **attrs)


def _distinct(nodes):
    seen, out = set(), []
    for name, kind, shape in nodes:
        if (kind, shape) not in seen:
            seen.add((kind, shape))
            out.append((name, kind, shape))
    return out


def _data(rng, kind, shape):
    if kind == "weight":
        return (rng.standard_normal(shape) * np.sqrt(2.0 / np.prod(shape[1:]))).astype(F)
    return rng.uniform(-1, 1, shape).astype(F)


@pytest.mark.parametrize("op_type,workload", [("Quantization_int8_V2", "resnet50_int8"),
                                              ("ClipGrad_Quantization_int8", "resnext101_clipgrad")])
def test_minmax_networks_every_distinct_node(op_type, workload):
    from b200quant.workloads import WORKLOADS
    rng = np.random.default_rng(5)
    nodes = _distinct(WORKLOADS[workload][0](2))
    assert len(nodes) > 20
    for name, kind, shape in nodes:
        is_w = kind == "weight"
        op, ref = _pair(op_type, quant_mode="minmax", is_weight=is_w, is_weight_perchannel=False, delay_quant=0,
                        ema_decay=0.99)
        aux_d, aux_r = [_dev(np.ones(1, F))], [np.ones(1, F)]
        for step in range(2):
            x, dy = _data(rng, kind, shape), rng.standard_normal(shape).astype(F)
            xd, yd, yr = _dev(x), _dev(np.zeros(shape, F)), np.zeros(shape, F)
            op.forward(True, ["write"], [xd], [yd], aux_d)
            ref.forward(True, ["write"], [x], [yr], aux_r)
            gd, gr = _dev(np.zeros(shape, F)), np.zeros(shape, F)
            op.backward(["write"], [_dev(dy)], [xd], [yd], [gd], aux_d)
            ref.backward(["write"], [dy], [x], [yr], [gr], aux_r)
            assert bits_equal(aux_d[0].cpu().numpy(), aux_r[0]), (name, step)
            assert bits_equal(yd.cpu().numpy(), yr), (name, step)
            assert bits_equal(gd.cpu().numpy(), gr), (name, step)


def test_mobilenet_gdrq_every_distinct_node():
    from b200quant.workloads import mobilenet_v1_nodes
    rng = np.random.default_rng(6)
    for name, kind, shape in _distinct(mobilenet_v1_nodes(2)):
        is_w = kind == "weight"
        for group_size in ((-1,) if len(shape) == 2 or (not is_w and shape[1] == 3) else (-1, 8 if not is_w else 4)):
            ch = shape[0] if is_w else shape[1]
            if group_size != -1 and ch % group_size:
                continue
            op, ref = _pair("GDRQ_PY", nbits=8, group_size=group_size, is_weight=is_w, lamda=0.001, delay_quant=0,
                            fix_alpha=False, ktimes=3)
            g = 1 if group_size == -1 else ch // group_size
            aux_d, aux_r = [_dev(np.ones(g, F))], [np.ones(g, F)]
            x, dy = _data(rng, kind, shape), rng.standard_normal(shape).astype(F)
            xd, yd, yr = _dev(x), _dev(np.zeros(shape, F)), np.zeros(shape, F)
            op.forward(True, ["write"], [xd], [yd], aux_d)
            ref.forward(True, ["write"], [x], [yr], aux_r)
            np.testing.assert_allclose(aux_d[0].cpu().numpy(), aux_r[0], rtol=1e-6, err_msg=name)
            assert bits_equal(aux_d[0].cpu().numpy(), aux_r[0]), (name, group_size)
            assert bits_equal(yd.cpu().numpy(), yr), (name, group_size)
            gd, gr = _dev(np.zeros(shape, F)), np.zeros(shape, F)
            op.backward(["write"], [_dev(dy)], [xd], [yd], [gd], aux_d)
            ref.backward(["write"], [dy], [x], [yr], [gr], aux_r)
            assert bits_equal(gd.cpu().numpy(), gr), (name, group_size)


def test_mobilenet_foldbn_layers():
    """fold-BN op on MobileNet-v1 layers (batch 2): quantised data, folded+quantised per-channel weight and folded bias
    bit-exact; convolution output to library tolerance."""
    rng = np.random.default_rng(7)
    # (cin, hw, cout, k, stride, pad, group): stem, depthwise / pointwise pairs of each stage, last pair
    layers = [(3, 224, 32, 3, 2, 1, 1), (32, 112, 32, 3, 1, 1, 32), (32, 112, 64, 1, 1, 0, 1), (64, 112, 64, 3, 2, 1, 64),
              (128, 56, 128, 3, 1, 1, 128), (256, 28, 512, 1, 1, 0, 1), (512, 14, 512, 3, 1, 1, 512),
              (512, 14, 1024, 1, 1, 0, 1), (1024, 7, 1024, 3, 1, 1, 1024), (1024, 7, 1024, 1, 1, 0, 1)]
    for cin, hw, cout, k, s, p, g in layers:
        attrs = dict(quant_mode="minmax", is_weight_perchannel=True, delay_quant=0, ema_decay=0.99, name="l",
                     num_filter=cout, num_group=g, kernel=(k, k), stride=(s, s), pad=(p, p), dilate=(1, 1),
                     no_bias=True, eps=1e-5, momentum=0.9, fix_gamma=False, quantize_flag=True)
        op, ref = _pair("GDRQ_Fold_BN", **attrs)
        oh = (hw + 2 * p - k) // s + 1
        ins = [rng.uniform(-1, 1, (2, cin, hw, hw)).astype(F), (rng.standard_normal((cout, cin // g, k, k)) * 0.2).astype(F),
               rng.standard_normal((2, cout, oh, oh)).astype(F), rng.uniform(0.5, 1.5, cout).astype(F),
               rng.standard_normal(cout).astype(F), rng.standard_normal(cout).astype(F), rng.uniform(0.5, 1.5, cout).astype(F)]
        aux_d = [_dev(np.ones(1, F)), _dev(np.ones(cout, F))]
        aux_r = [np.ones(1, F), np.ones(cout, F)]
        yd, yr = _dev(np.zeros((2, cout, oh, oh), F)), np.zeros((2, cout, oh, oh), F)
        op.forward(True, ["write"], [_dev(a) for a in ins], [yd], aux_d)
        ref.forward(True, ["write"], [a.copy() for a in ins], [yr], aux_r)
        tag = (cin, hw, cout, k, g)
        assert bits_equal(aux_d[0].cpu().numpy(), aux_r[0]) and bits_equal(aux_d[1].cpu().numpy(), aux_r[1]), tag
        assert bits_equal(op.data_q.cpu().numpy(), ref.data_q), tag
        assert bits_equal(op.weight_q.cpu().numpy(), ref.weight_q), tag
        assert bits_equal(op.bias.cpu().numpy(), ref.bias), tag
        np.testing.assert_allclose(yd.cpu().numpy(), yr, rtol=2e-4, atol=2e-4)
        grads = [_dev(np.full(a.shape, 3, F)) for a in ins]
        dy = rng.standard_normal(yr.shape).astype(F)
        op.backward(["write"] * 7, [_dev(dy)], [_dev(a) for a in ins], [yd], grads, aux_d)
        for i, gten in enumerate(grads):
            want = dy if i == 2 else np.zeros(ins[i].shape, F)
            assert bits_equal(gten.cpu().numpy(), want), (tag, i)
