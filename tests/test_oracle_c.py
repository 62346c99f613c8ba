"""The C/OpenMP restatement (the timed CPU baseline) must agree with the NumPy oracle bit for bit."""
import numpy as np
import pytest

from oracle import c_oracle as co
from oracle import quant_oracle as qo
from tests.golden_util import bits_equal

F = np.float32


@pytest.mark.parametrize("variant,name", [(0, "Quantization_int8_V2"), (1, "ClipGrad_Quantization_int8")])
@pytest.mark.parametrize("is_weight,per_channel,shape", [(False, False, (4, 6, 9, 9)), (True, False, (16, 8, 3, 3)),
                                                         (True, True, (16, 8, 3, 3)), (True, True, (32, 1, 3, 3))])
def test_minmax_ops(variant, name, is_weight, per_channel, shape):
    rng = np.random.default_rng(21)
    ref = qo.create(name, quant_mode="minmax", is_weight=str(is_weight), is_weight_perchannel=str(per_channel))
    naux = shape[0] if per_channel else 1
    aux_c, aux_r = np.ones(naux, F), np.ones(naux, F)
    init = True
    for step, train in enumerate([True, True, False]):
        x = (rng.standard_normal(shape) * (1 + step)).astype(F)
        dy = rng.standard_normal(shape).astype(F)
        yc, yr = np.zeros(shape, F), np.zeros(shape, F)
        co.minmax_quant_fwd(variant, x, yc, aux_c, is_weight, per_channel, train, init and train and not is_weight, 0.99)
        if train:
            init = False
        ref.forward(train, ["write"], [x], [yr], [aux_r])
        assert bits_equal(aux_c, aux_r) and bits_equal(yc, yr)
        gc, gr = np.zeros(shape, F), np.zeros(shape, F)
        if variant == 1 and not is_weight:
            co.clipgrad_bwd(x, dy, gc, aux_c)
        else:
            co.ste_bwd(dy, gc)
        ref.backward(["write"], [dy], [x], [yr], [gr], [aux_r])
        assert bits_equal(gc, gr)


@pytest.mark.parametrize("is_weight", [True, False])
def test_gdrq_pertensor(is_weight):
    rng = np.random.default_rng(22)
    ref = qo.create("GDRQ_PY", nbits="8", group_size="-1", is_weight=str(is_weight), lamda="0.001", delay_quant="1",
                    fix_alpha="False", ktimes="3")
    a_c, a_r = np.ones(1, F), np.ones(1, F)
    for step in range(3):
        x = rng.standard_normal((3, 8, 7, 7)).astype(F)
        dy = rng.standard_normal(x.shape).astype(F)
        yc, yr = np.zeros_like(x), np.zeros_like(x)
        co.gdrq_fwd(x, yc, a_c, is_weight, False, step >= 1, 255, 3, 0.001)
        ref.forward(True, ["write"], [x], [yr], [a_r])
        assert bits_equal(a_c, a_r) and bits_equal(yc, yr)
        gc, gr = np.zeros_like(x), np.zeros_like(x)
        if is_weight:
            co.ste_bwd(dy, gc)
        else:
            co.gdrq_bwd(x, dy, gc, a_c)
        ref.backward(["write"], [dy], [x], [yr], [gr], [a_r])
        assert bits_equal(gc, gr)


@pytest.mark.parametrize("per_channel", [True, False])
def test_foldbn_pieces(per_channel):
    rng = np.random.default_rng(23)
    cout, cin, k = 12, 5, 3
    ref = qo.create("GDRQ_Fold_BN", quant_mode="minmax", is_weight_perchannel=str(per_channel), num_filter=str(cout),
                    num_group="1", kernel="(3,3)", stride="(1,1)", pad="(1,1)")
    aux_r = [np.ones(1, F), np.ones(cout if per_channel else 1, F)]
    aux_c = [a.copy() for a in aux_r]
    for step in range(2):
        ins = [rng.uniform(-1, 1, (2, cin, 6, 6)).astype(F), (rng.standard_normal((cout, cin, k, k)) * 0.3).astype(F),
               rng.standard_normal((2, cout, 6, 6)).astype(F), rng.uniform(0.5, 1.5, cout).astype(F),
               rng.standard_normal(cout).astype(F), rng.standard_normal(cout).astype(F),
               rng.uniform(0.5, 1.5, cout).astype(F)]
        out = np.zeros((2, cout, 6, 6), F)
        ref.forward(True, ["write"], ins, [out], aux_r)
        xq, wq, bias = np.zeros_like(ins[0]), np.zeros_like(ins[1]), np.zeros(cout, F)
        co.foldbn_data_fwd(ins[0], xq, aux_c[0], step == 0, 0.99)
        co.foldbn_weight_fwd(ins[1], wq, bias, aux_c[1], ins[3], ins[4], ins[5], ins[6], 1e-5, per_channel, True, True)
        assert bits_equal(aux_c[0], aux_r[0]) and bits_equal(aux_c[1], aux_r[1])
        assert bits_equal(xq, ref.data_q) and bits_equal(wq, ref.weight_q) and bits_equal(bias, ref.bias)


def _same_or_nan(a, b):
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb])


@pytest.mark.parametrize("kind", ["zeros", "nan", "inf", "denormal", "outlier"])
@pytest.mark.parametrize("variant,name", [(0, "Quantization_int8_V2"), (1, "ClipGrad_Quantization_int8")])
@pytest.mark.parametrize("is_weight,per_channel", [(False, False), (True, False), (True, True)])
def test_minmax_ops_edge_inputs(variant, name, is_weight, per_channel, kind):
    """all-zero / NaN / Inf / denormal / outlier inputs: the C restatement used for the full-size GPU
    parity tests must carry NaN and Inf exactly like the NumPy oracle (max|x| of a tensor holding NaN is NaN)."""
    rng = np.random.default_rng(23)
    shape = (6, 5, 3, 3)
    x = rng.standard_normal(shape).astype(F)
    flat = x.reshape(-1)
    if kind == "zeros":
        flat[:] = 0
    elif kind == "nan":
        flat[100] = np.nan
    elif kind == "inf":
        flat[100] = -np.inf
    elif kind == "denormal":
        flat *= F(1e-41)
    else:
        flat[100] = F(3e30)
    ref = qo.create(name, quant_mode="minmax", is_weight=str(is_weight), is_weight_perchannel=str(per_channel))
    naux = shape[0] if per_channel else 1
    aux_c, aux_r = np.ones(naux, F), np.ones(naux, F)
    yc, yr = np.zeros(shape, F), np.zeros(shape, F)
    with np.errstate(all="ignore"):
        co.minmax_quant_fwd(variant, x, yc, aux_c, is_weight, per_channel, True, not is_weight, 0.99)
        ref.forward(True, ["write"], [x], [yr], [aux_r])
    assert _same_or_nan(aux_c, aux_r) and _same_or_nan(yc, yr)
