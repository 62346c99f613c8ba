"""Drop-in for ``core/operator/QIL.py`` (op_type ``QIL_PY``); QIL_V2.py / QIL_V3.py build on it.

Quantization-interval learning: out = sign(x)([|x| > cp] + (a|x| + b)[pp <= |x| <= cp]) rounded to 2^nbits-1
levels; the variants differ in how (pp, cp, a, b) derive from the two learnable scalars.  The reference clamps
pruning_point >= 0 and clipping_point <= 1 in place with host reads and asserts pp < cp on the host
(QIL.py:51-57); the clamp is done on the device here and the assert is dropped (it would force a device sync)."""
from .. import _kernels as K
from ..operator import CustomOp, CustomOpProp, py_bool, register


class QIL_PY(CustomOp):
    VARIANT = 1

    def __init__(self, is_weight, fix_gamma, nbits):
        self.is_weight = is_weight
        self.fix_gamma = fix_gamma
        self.nbits = nbits
        self.QUANT_LEVEL = 2 ** self.nbits - 1
        self.count = 0
        self.quantized_type = "ste"

    def forward(self, is_train, req, in_data, out_data, aux):
        assert len(in_data) == 4, "QIL takes data, two interval parameters and gamma"
        K.qil_fwd(self.VARIANT, in_data[0], out_data[0], in_data[1], in_data[2], self.QUANT_LEVEL, req[0])

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        assert len(req) >= 3
        assert self.fix_gamma == True, "currently only support fix gamma"  # noqa: E712
        K.qil_bwd(self.VARIANT, in_data[0], out_grad[0], in_grad[0], in_data[1], in_data[2], in_grad[1], in_grad[2],
                  req[0], req[1], req[2])


class _QILProp(CustomOpProp):
    OP = None
    ARGS = None

    def __init__(self, is_weight="False", fix_gamma="True", nbits="4"):
        self.is_weight = py_bool(is_weight)
        self.fix_gamma = py_bool(fix_gamma)
        self.nbits = int(nbits)
        super(_QILProp, self).__init__(True)

    def list_arguments(self):
        return list(self.ARGS)

    def infer_shape(self, in_shape):
        shape = in_shape[0]
        return [shape, [1], [1], [1]], [shape], []

    def create_operator(self, ctx, shapes, dtypes):
        return self.OP(self.is_weight, self.fix_gamma, self.nbits)


@register("QIL_PY")
class QIL_PYProp(_QILProp):
    OP = QIL_PY
    ARGS = ["data", "pruning_point", "clipping_point", "gamma"]
