"""The drop-in boundary on the MXNet side, without MXNet: (1) every operator's Prop has the contracts the reference's
Props have (names of arguments / outputs / auxiliary states, infer_shape) -- compared with what the reference classes
reported when the golden fixtures were generated; (2) when a module named ``mxnet`` is importable the classes subclass
ITS ``mx.operator.CustomOp`` / ``CustomOpProp`` and register through ITS ``mx.operator.register`` under the reference's
op_type strings (symbol/quant_ops.py:44, clip_grad_quantization_int8.py:70, fold_bn_v1_gdrq.py:131, GDRQ.py:154,210,
PACT.py, WNQ.py, QIL*.py), which is what makes ``mx.sym.Custom(op_type=...)`` find them."""
import os
import subprocess
import sys

import pytest

from tests.golden_util import CASES, MANIFEST

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

OP_TYPES = ["Quantization_int8_V2", "ClipGrad_Quantization_int8", "GDRQ_Fold_BN", "GDRQ_PY", "CLIP_RELU_PY", "DoReFa_PY",
            "PACT_PY", "PACT_V2_PY", "QUANT_STE_PY", "WNQ_PY", "QIL_PY", "QIL_V2_PY", "QIL_V3_PY"]


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_prop_contracts_equal_the_reference(case):
    import b200quant
    prop = b200quant.get_prop(case["op_type"])(**case["attrs"])
    man = MANIFEST[case["id"]]
    assert list(prop.list_arguments()) == man["list_arguments"]
    assert list(prop.list_outputs()) == man["list_outputs"]
    assert list(prop.list_auxiliary_states()) == man["list_auxiliary_states"]
    in_shapes, out_shapes, aux_shapes = prop.infer_shape(man["in_shapes"])
    assert [list(s) for s in in_shapes] == man["in_shapes"]
    assert [list(s) for s in out_shapes] == man["out_shapes"]
    assert [list(s) for s in aux_shapes] == man["aux_shapes"]


def test_every_op_type_of_the_reference_is_registered():
    import b200quant
    from b200quant.operator import REGISTRY
    assert sorted(REGISTRY) == sorted(OP_TYPES)


_FAKE_MXNET = r'''
import sys, types
mx = types.ModuleType("mxnet")
op = types.ModuleType("mxnet.operator")
op.registered = {}
class CustomOp(object):
    def assign(self, dst, req, src): raise NotImplementedError
class CustomOpProp(object):
    def __init__(self, need_top_grad=True): self.need_top_grad_ = need_top_grad
def register(name):
    def deco(cls):
        assert issubclass(cls, CustomOpProp), cls
        op.registered[name] = cls
        return cls
    return deco
op.CustomOp, op.CustomOpProp, op.register = CustomOp, CustomOpProp, register
mx.operator = op
sys.modules["mxnet"], sys.modules["mxnet.operator"] = mx, op
sys.path.insert(0, %r)
import b200quant
from b200quant import operator as bop
assert bop.HAVE_MXNET and bop.CustomOp is CustomOp and bop.CustomOpProp is CustomOpProp
for name, cls in op.registered.items():
    prop = cls(**%r.get(name, {}))
    inst = prop.create_operator(None, None, None)
    assert isinstance(inst, CustomOp), name
    # the engine cannot see the library's launches: forward / backward synchronise before returning
    assert getattr(type(inst).forward, "__b2q_syncs__", False) and getattr(type(inst).backward, "__b2q_syncs__", False), name
print(",".join(sorted(op.registered)))
'''


def test_classes_register_with_an_importable_mxnet():
    attrs = {"Quantization_int8_V2": {"quant_mode": "minmax", "is_weight": "True"},
             "ClipGrad_Quantization_int8": {"quant_mode": "minmax", "is_weight": "False"},
             "GDRQ_Fold_BN": {"quant_mode": "minmax", "num_filter": "8", "num_group": "1"}}
    code = _FAKE_MXNET % (ROOT, attrs)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    assert out.stdout.strip().split(",") == sorted(OP_TYPES)
