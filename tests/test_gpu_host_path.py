"""Host-buffer path (bench.py "e2e"): CPU tensors handed to the operators are staged through device memory by the
library (three-stream pipeline over a staging ring); results must equal the oracle bit for bit, for many tensors in
flight at once, of very different sizes, pinned or pageable."""
import numpy as np
import pytest

from oracle import quant_oracle as qo
from tests.golden_util import bits_equal

pytestmark = pytest.mark.gpu
F = np.float32


@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("op_type", ["Quantization_int8_V2", "ClipGrad_Quantization_int8"])
def test_host_tensors_many_in_flight(op_type, pinned):
    import torch
    import b200quant
    from b200quant import _kernels as K
    rng = np.random.default_rng(5)
    shapes = [(64, 3, 7, 7), (8, 64, 56, 56), (3,), (256, 2048), (16, 128, 28, 28), (1000, 2048), (2, 3, 224, 224),
              (64, 64, 1, 1), (4, 256, 56, 56), (7, 5), (512, 512, 3, 3)] * 2
    jobs = []
    for i, shape in enumerate(shapes):
        is_w = len(shape) == 4 and shape[0] in (64, 512) or shape == (1000, 2048)
        pc = bool(is_w and i % 2)
        op = b200quant.get_prop(op_type)(quant_mode="minmax", is_weight=str(is_w), is_weight_perchannel=str(pc)) \
            .create_operator(None, None, None)
        ref = qo.create(op_type, quant_mode="minmax", is_weight=str(is_w), is_weight_perchannel=str(pc))
        x = (rng.standard_normal(shape) * (0.1 + i)).astype(F)
        dy = rng.standard_normal(shape).astype(F)
        naux = shape[0] if pc else 1
        mk = (lambda a: torch.from_numpy(a.copy()).pin_memory()) if pinned else (lambda a: torch.from_numpy(a.copy()))
        jobs.append(dict(op=op, ref=ref, x=x, dy=dy, hx=mk(x), hy=mk(np.zeros(shape, F)), hdy=mk(dy),
                         hdx=mk(np.zeros(shape, F)), haux=mk(np.ones(naux, F)), aux_r=np.ones(naux, F), is_w=is_w))
    for j in jobs:                                   # everything enqueued before anything is waited for
        j["op"].forward(True, ["write"], [j["hx"]], [j["hy"]], [j["haux"]])
    K.host_sync()
    for j in jobs:
        j["op"].backward(["write"], [j["hdy"]], [j["hx"]], [j["hy"]], [j["hdx"]], [j["haux"]])
    K.host_sync()
    for i, j in enumerate(jobs):
        yr, gr = np.zeros_like(j["x"]), np.zeros_like(j["x"])
        j["ref"].forward(True, ["write"], [j["x"]], [yr], [j["aux_r"]])
        j["ref"].backward(["write"], [j["dy"]], [j["x"]], [yr], [gr], [j["aux_r"]])
        assert bits_equal(j["haux"].numpy(), j["aux_r"]), i
        assert bits_equal(j["hy"].numpy(), yr), i
        assert bits_equal(j["hdx"].numpy(), gr), i


def test_numpy_arrays_are_accepted_zero_copy():
    import b200quant
    from b200quant import _kernels as K
    rng = np.random.default_rng(6)
    x = rng.standard_normal((32, 8, 16, 16)).astype(F)
    y, aux = np.zeros_like(x), np.ones(1, F)
    op = b200quant.get_prop("Quantization_int8_V2")(quant_mode="minmax", is_weight="False").create_operator(None, None, None)
    ref = qo.create("Quantization_int8_V2", quant_mode="minmax", is_weight="False")
    op.forward(True, ["write"], [x], [y], [aux])
    K.host_sync()
    yr, ar = np.zeros_like(x), np.ones(1, F)
    ref.forward(True, ["write"], [x.copy()], [yr], [ar])
    assert bits_equal(aux, ar) and bits_equal(y, yr)


@pytest.mark.parametrize("n", [1, 15, 16, 17, 63, 64, 65, 255, 256, 257, 4099, (8 << 20) // 4 + 13, 3 * (8 << 20) // 4 + 5])
def test_host_ste_copy_any_alignment(n):
    """The straight-through backward of a host caller (quant_ops.py:42-43) is a host-to-host streaming copy (64-byte
    non-temporal stores with a head/tail split): every source / destination misalignment, sizes around the vector
    width and the 8 MB job size, and nothing outside [dst, dst + n) is touched."""
    from b200quant import _kernels as K
    rng = np.random.default_rng(n)
    for so in (0, 1, 3):
        for do in (0, 1, 5, 15):
            src = rng.standard_normal(n + 32).astype(F)
            dst = np.full(n + 48, 7.0, F)
            K.ste_bwd(src[so:so + n], dst[do:do + n], "write")
            K.host_sync()
            assert bits_equal(dst[do:do + n], src[so:so + n])
            assert (dst[:do] == 7.0).all() and (dst[do + n:] == 7.0).all()
