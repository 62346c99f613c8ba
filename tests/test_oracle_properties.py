"""Property tests of the oracle's MXNet numerics (hypothesis): roundf semantics against libm, exact ties, the
known-answer vectors of SURVEY.md F8, and C-restatement == NumPy oracle on random shapes."""
import ctypes
import ctypes.util

import numpy as np
from hypothesis import given, settings, strategies as st
from hypothesis.extra import numpy as hnp

from oracle import c_oracle as co
from oracle import quant_oracle as qo
from tests.golden_util import bits_equal

F = np.float32
_libm = ctypes.CDLL(ctypes.util.find_library("m"))
_libm.roundf.restype = ctypes.c_float
_libm.roundf.argtypes = [ctypes.c_float]


@given(hnp.arrays(np.float32, st.integers(1, 200), elements=st.floats(-1e6, 1e6, width=32)))
@settings(max_examples=200, deadline=None)
def test_mx_round_is_libm_roundf(x):
    want = np.array([_libm.roundf(float(v)) for v in x], dtype=F)
    assert bits_equal(qo.mx_round(x), want)


def test_mx_round_known_answers():
    x = np.array([0.5, -0.5, 1.5, 2.5, -2.5, 0.49999997, -0.49999997, 8388607.5, -0.3, 0.0, -0.0, 126.5, -126.5], F)
    want = np.array([1, -1, 2, 3, -3, 0, -0.0, 8388608, -0.0, 0, -0.0, 127, -127], F)
    assert bits_equal(qo.mx_round(x), want)
    assert not np.array_equal(np.round(x), want)          # np.round (half-even) is NOT what MXNet does


@given(st.integers(-126, 126), st.sampled_from([1.0, 0.5, 127.0 / 8, 3.0]))
def test_exact_ties_round_away(k, t):
    thr = F(t)
    q = qo.mx_div(thr, F(127))
    x = F((F(k) + F(0.5)) * q)
    y, code = qo.qdq(np.array([x], F), q)
    if float(qo.mx_div(x, q)) == k + 0.5:                  # the tie survived the float32 division
        assert float(code[0]) == (k + 1 if k >= 0 else k)


@given(st.integers(1, 6), st.integers(1, 5), st.integers(1, 4), st.integers(1, 4), st.booleans(), st.booleans(),
       st.integers(0, 2 ** 31 - 1))
@settings(max_examples=40, deadline=None)
def test_c_restatement_equals_numpy_oracle(c, i, h, w, is_weight, per_channel, seed):
    rng = np.random.default_rng(seed)
    shape = (c, i, h, w)
    per_channel = per_channel and is_weight
    for variant, name in ((0, "Quantization_int8_V2"), (1, "ClipGrad_Quantization_int8")):
        ref = qo.create(name, quant_mode="minmax", is_weight=str(is_weight), is_weight_perchannel=str(per_channel))
        naux = c if per_channel else 1
        aux_c, aux_r = np.full(naux, 0.7, F), np.full(naux, 0.7, F)
        x = (rng.standard_normal(shape) * rng.uniform(0.01, 10)).astype(F)
        x.flat[0] = 0.37                                   # never an all-zero tensor (q = 0 gives NaN on both sides)
        yc, yr = np.zeros(shape, F), np.zeros(shape, F)
        co.minmax_quant_fwd(variant, x, yc, aux_c, is_weight, per_channel, True, not is_weight, 0.99)
        ref.forward(True, ["write"], [x], [yr], [aux_r])
        assert bits_equal(aux_c, aux_r) and bits_equal(yc, yr)
