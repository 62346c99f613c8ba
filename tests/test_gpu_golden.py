"""CUDA operators (through the CustomOp protocol -> ctypes -> libb2q.so) against the golden fixtures that were
produced by the reference's own op classes (tests/golden/generate.py)."""
import numpy as np
import pytest

from tests.golden_util import CASES, MANIFEST, check_against_golden, load
from tests.golden.driver import drive

pytestmark = pytest.mark.gpu


def _mk(case):
    import torch
    import b200quant
    prop = b200quant.get_prop(case["op_type"])(**case["attrs"])
    man = MANIFEST[case["id"]]
    assert list(prop.list_arguments()) == man["list_arguments"]
    assert list(prop.list_outputs()) == man["list_outputs"]
    assert list(prop.list_auxiliary_states()) == man["list_auxiliary_states"]
    _, out_shapes, aux_shapes = prop.infer_shape(man["in_shapes"])
    assert [list(s) for s in out_shapes] == man["out_shapes"]
    assert [list(s) for s in aux_shapes] == man["aux_shapes"]
    op = prop.create_operator(None, None, None)
    to_arr = lambda a: torch.from_numpy(np.array(a, dtype=np.float32)).cuda()
    to_np = lambda t: t.detach().cpu().numpy()
    return op, to_arr, to_np


_DOREFA = [c for c in CASES if c["op_type"] == "DoReFa_PY"]
_REST = [c for c in CASES if c["op_type"] != "DoReFa_PY"]


@pytest.mark.parametrize("case", _REST, ids=[c["id"] for c in _REST])
def test_cuda_op_matches_reference_fixture(case):
    op, to_arr, to_np = _mk(case)
    check_against_golden(case, op, to_arr, to_np)


@pytest.mark.parametrize("case", _DOREFA, ids=[c["id"] for c in _DOREFA])
def test_dorefa_matches_up_to_tanh_ulps(case):
    """tanhf differs by ulps between libraries; a code may flip only where L*o sits on a rounding boundary."""
    op, to_arr, to_np = _mk(case)
    fx = load(case["id"])
    man = MANIFEST[case["id"]]
    got = drive(case, fx, op, [tuple(s) for s in man["aux_shapes"]], to_arr, to_np)
    L = 2 ** int(case["attrs"]["nbits"]) - 1
    for k in man["keys"]:
        ref, val = fx[k], got[k]
        if k.endswith("_out"):
            diff = np.abs(val - ref)
            assert np.all(diff <= 2.0 / L + 1e-6)
            assert np.mean(diff > 1e-6) < 0.02
        elif "_ig" in k:
            np.testing.assert_allclose(val, ref, rtol=1e-4, atol=1e-5)
        else:
            np.testing.assert_allclose(val, ref, rtol=1e-6, atol=1e-7)
