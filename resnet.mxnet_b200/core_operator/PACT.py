"""Drop-in for ``core/operator/PACT.py``: op_types DoReFa_PY, PACT_PY, PACT_V2_PY, QUANT_STE_PY.

The reference records an autograd graph in forward and replays it in backward (PACT.py:46-49,124-125,188-193);
here the same derivatives are evaluated by one fused kernel per direction (elementwise gradient + the
block-reduced scalar gradient of the learnable threshold)."""
from .. import _kernels as K
from .. import _lib
from ..operator import CustomOp, CustomOpProp, py_literal, register


def _scratch_like(ref, n=1):
    return K.scratch_like(ref, n)


class DoReFa_PY(CustomOp):
    """PACT.py:30-77: w -> tanh(w) / (2 max|tanh(w)|) + 0.5 -> k-bit uniform -> 2x-1."""

    def __init__(self, nbits):
        self.nbits = nbits
        self.data = None
        self.output = None
        self._vmax = None

    def forward(self, is_train, req, in_data, out_data, aux):
        self._vmax = _scratch_like(in_data[0])
        K.dorefa_fwd(in_data[0], out_data[0], self._vmax, 2 ** self.nbits - 1, req[0])

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        K.dorefa_bwd(in_data[0], out_grad[0], in_grad[0], self._vmax, req[0])


class _NbitsProp(CustomOpProp):
    OP = None
    ARGS = ["data"]

    def __init__(self, nbits="8"):
        self.nbits = py_literal(nbits)
        super(_NbitsProp, self).__init__(True)

    def list_arguments(self):
        return list(self.ARGS)

    def infer_shape(self, in_shape):
        shape = in_shape[0]
        return [shape] + [[1]] * (len(self.ARGS) - 1), [shape], []

    def create_operator(self, ctx, shapes, dtypes):
        return self.OP(self.nbits)


@register("DoReFa_PY")
class DoReFa_PYProp(_NbitsProp):
    OP = DoReFa_PY


class PACT_PY(CustomOp):
    """PACT.py:102-144: one-sided learnable clip gamma (an *argument*, not aux), 2^nbits-1 levels."""
    TWO_SIDED = False

    def __init__(self, nbits):
        self.nbits = nbits
        self.QUANT_LEVEL = 2 ** self.nbits - 1
        self.count = 0

    def forward(self, is_train, req, in_data, out_data, aux):
        assert len(in_data) == 2, "the input must be 2 in PACT: data and gamma"
        K.qdq(in_data[0], out_data[0], in_data[1], self.QUANT_LEVEL,
              _lib.CLIP_WHERE_LT if self.TWO_SIDED else _lib.CLIP_PACT, req[0])

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        K.pact_bwd(in_data[0], out_grad[0], in_grad[0], in_grad[1], in_data[1], self.TWO_SIDED, req[0], req[1])


@register("PACT_PY")
class PACT_PYProp(_NbitsProp):
    OP = PACT_PY
    ARGS = ["data", "gamma"]


class PACT_V2_PY(PACT_PY):
    """PACT.py:169-203: two-sided clip where(|x| < gamma, x, gamma*sign(x))."""
    TWO_SIDED = True


@register("PACT_V2_PY")
class PACT_V2_PYProp(_NbitsProp):
    OP = PACT_V2_PY
    ARGS = ["data", "gamma"]


class QUANT_STE_PY(CustomOp):
    """PACT.py:238-252: absmax scaling to 2^(nbits-1)-1 levels, straight-through backward."""

    def __init__(self, nbits):
        self.nbits = nbits
        self.QUANT_LEVEL = 2 ** (self.nbits - 1) - 1
        self.count = 0

    def forward(self, is_train, req, in_data, out_data, aux):
        m = _scratch_like(in_data[0])
        K.absmax(in_data[0], m)
        K.qdq(in_data[0], out_data[0], m, self.QUANT_LEVEL, _lib.CLIP_NONE, req[0])

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        K.ste_bwd(out_grad[0], in_grad[0], req[0])


@register("QUANT_STE_PY")
class QUANT_STE_PYProp(_NbitsProp):
    OP = QUANT_STE_PY
