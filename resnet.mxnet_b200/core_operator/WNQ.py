"""Drop-in for ``core/operator/WNQ.py``: op_type ``WNQ_PY`` (weight-normalised quantisation, non-STE backward)."""
from .. import _kernels as K
from ..operator import CustomOp, CustomOpProp, py_bool, register


class WNQ_PY(CustomOp):
    """WNQ.py:46-85: y = round((w/m) L)/L * m with m = max|w| per tensor / per out-channel; the backward routes
    -(sum dy*w*[|w|!=m])/m to the max element(s) and dy to the others."""

    def __init__(self, nbits, is_perchannel):
        self.nbits = nbits
        self.is_perchannel = is_perchannel
        self.QUANT_LEVEL = 2 ** self.nbits - 1

    def forward(self, is_train, req, in_data, out_data, aux):
        K.wnq_fwd(in_data[0], out_data[0], self.is_perchannel, self.QUANT_LEVEL, req[0])

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        K.wnq_bwd(in_data[0], out_grad[0], in_grad[0], self.is_perchannel, req[0])


@register("WNQ_PY")
class WNQ_PYProp(CustomOpProp):
    def __init__(self, nbits=4, is_perchannel=False):
        self.nbits = int(nbits)
        self.is_perchannel = py_bool(is_perchannel)
        super(WNQ_PYProp, self).__init__(True)

    def infer_shape(self, in_shape):
        return [in_shape[0]], [in_shape[0]], []

    def create_operator(self, ctx, shapes, dtypes):
        return WNQ_PY(self.nbits, self.is_perchannel)
