"""Drop-in for ``symbol/clip_grad_quantization_int8.py``: op_type ``ClipGrad_Quantization_int8``.

Differences from Quantization_int8_V2 that are kept (clip_grad_quantization_int8.py:14-67): aux ``minmax`` is
the source of truth for the scale (weights quantise from aux, so eval uses the stored max), the activation
EMA is initialised from the first training batch, activations are clipped to +-aux before rounding and written
with ``[:]=`` regardless of ``req``, and the activation backward masks the gradient to ``-T < x < T`` (strict)
with ``T`` = aux after this step's update.  No host round trip: the reference's two ``asnumpy()`` syncs
(:49-50) are replaced by the kernels reading aux from device memory.
"""
from . import _kernels as K
from .operator import register
from .quant_ops import Quantization_int8, _MinMaxProp


class ClipGrad_Quantization_int8(Quantization_int8):
    """symbol/clip_grad_quantization_int8.py:5-67."""
    VARIANT = 1

    def forward(self, is_train, req, in_data, out_data, aux):
        if is_train and self.delay_quant > 0:      # :15-18
            self.assign(out_data[0], req[0], in_data[0])
            self.delay_quant -= 1
            return
        first = bool(self.init) and bool(is_train) and not self.is_weight
        self._quantize(is_train, req[0], in_data[0], out_data[0], aux[0], first)
        if first:
            self.init = False                       # :42-44

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        if self.is_weight:
            K.ste_bwd(out_grad[0], in_grad[0], req[0])                    # :58-59
        else:
            K.clipgrad_bwd(in_data[0], out_grad[0], in_grad[0], aux[0])   # :61-67 (ignores req)


@register("ClipGrad_Quantization_int8")
class ClipGradQuantizationInt8Prop(_MinMaxProp):
    OP = ClipGrad_Quantization_int8
