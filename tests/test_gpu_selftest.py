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


def test_tanhf_is_odd_and_monotonic_over_every_positive_float():
    """DoReFa_PY's max|tanh(w)| is computed as tanhf(max|w|); that is the same float only if tanhf is odd and monotonic
    non-decreasing, which is checked here for all 2^31 finite positive floats."""
    assert _selftest(2) == 0


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
