"""Drop-in for the reference's ``symbol/quant_ops.py``: op_type ``Quantization_int8_V2``.

Same class names, op_type, attribute names / defaults / string parsing, argument, output and aux names and
shapes as the reference (symbol/quant_ops.py:3-72); the arithmetic runs in libb2q.so:

    weight      one fused max|w| reduction (per tensor or per out-channel) + one QDQ sweep
    activation  one fused max|x| reduction whose last block applies the EMA update of ``minmax`` + one QDQ sweep
                (no clip: codes may exceed +-127 exactly as in the reference, quant_ops.py:39-40)
    backward    straight-through copy honouring ``req``
"""
from . import _kernels as K
from .operator import CustomOp, CustomOpProp, py_bool, register

from ._mx import mx   # resolved at call time: the symbol builders need MXNet, the operators do not


class Quantization_int8(CustomOp):
    """symbol/quant_ops.py:3-42."""
    VARIANT = 0

    def __init__(self, quant_mode, is_weight, is_weight_perchannel, delay_quant, ema_decay):
        self.quant_mode = quant_mode
        self.is_weight = is_weight
        self.is_weight_perchannel = is_weight_perchannel
        self.delay_quant = delay_quant
        self.ema_decay = ema_decay
        self.QUANT_LEVEL = 127
        self.init = True
        self.sync = None     # optional cross-rank threshold exchange through NCCL (b200quant.dist.ThresholdSync)
        self.peer = None     # optional fused peer-memory exchange (b200quant.dist.PeerThresholdExchange)
        self._stat = None

    def _quantize(self, is_train, req, x, y, aux, first):
        """One fused call; or, when a cross-rank sync is attached to a training activation node:
        reduce -> allreduce(max) -> update + QDQ, so every rank applies the same threshold."""
        if self.peer is not None and is_train and not self.is_weight and req in ("write", "inplace"):
            self.peer.quantize(self.VARIANT, x, y, aux, first, self.ema_decay)
            return
        if self.peer is not None and self.sync is None and is_train and not self.is_weight:
            # a peer exchange is attached but this call is outside its fused case (req='add'): the threshold must still
            # be the max over ranks, so go through the NCCL exchange instead of silently using the local statistic
            from .dist import ThresholdSync
            self.sync = ThresholdSync(getattr(self.peer, "group", None))
        if self.sync is not None and is_train and not self.is_weight:
            if self._stat is None:
                self._stat = K.scratch_like(x, 1)
            K.minmax_quant_stat(x, self._stat, False)
            self.sync(self._stat)
            K.minmax_quant_finish(self.VARIANT, x, y, aux, self._stat, False, False, True, first, self.ema_decay, req)
        else:
            K.minmax_quant_fwd(self.VARIANT, x, y, aux, self.is_weight, self.is_weight_perchannel, is_train, first,
                               self.ema_decay, req)

    def forward(self, is_train, req, in_data, out_data, aux):
        if is_train and self.delay_quant > 0:      # :13-16 warm-up: pass through, count down
            self.assign(out_data[0], req[0], in_data[0])
            self.delay_quant -= 1
            return
        self._quantize(is_train, req[0], in_data[0], out_data[0], aux[0], False)

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        K.ste_bwd(out_grad[0], in_grad[0], req[0])  # :41-42


class _MinMaxProp(CustomOpProp):
    """Shared Prop of the two minmax operators (quant_ops.py:44-72, clip_grad_quantization_int8.py:70-97)."""
    OP = None

    def __init__(self, quant_mode, is_weight, is_weight_perchannel=False, delay_quant=0, ema_decay=0.99):
        self.quant_mode = str(quant_mode)
        self.delay_quant = int(delay_quant)
        self.ema_decay = float(ema_decay)
        self.is_weight = py_bool(is_weight)
        self.is_weight_perchannel = py_bool(is_weight_perchannel)
        super(_MinMaxProp, self).__init__(True)

    def list_arguments(self):
        return ["data"]

    def list_outputs(self):
        return ["output"]

    def list_auxiliary_states(self):
        return ["minmax"]

    def infer_shape(self, in_shape):
        shape = in_shape[0]
        aux_shape = [shape[0]] if (self.is_weight_perchannel and self.is_weight) else [1]
        return [shape], [shape], [aux_shape]

    def infer_type(self, in_type):
        return in_type, in_type, in_type

    def create_operator(self, ctx, shapes, dtypes):
        return self.OP(self.quant_mode, self.is_weight, self.is_weight_perchannel, self.delay_quant, self.ema_decay)


@register("Quantization_int8_V2")
class QuantizationInt8Prop(_MinMaxProp):
    OP = Quantization_int8


# ---- symbol builders (quant_ops.py:75-121); need MXNet's symbolic API --------------------------------------
def _need_mx():
    return mx.module()


def get_sym_output_channel(name, sym, data_shape=(1, 3, 224, 224)):
    _need_mx()
    _, out_shapes, _ = sym.infer_shape(data=data_shape)
    assert len(out_shapes) == 1, "the output of sym is not equal to 1"
    return out_shapes[0][1]


def _qnode(data, name, is_weight, quant_mod, delay_quant, is_weight_perchannel, op_type="Quantization_int8_V2"):
    return mx.sym.Custom(data=data, name=name, quant_mode=quant_mod, is_weight=is_weight,
                         is_weight_perchannel=is_weight_perchannel, ema_decay=0.99, delay_quant=delay_quant,
                         op_type=op_type)


def quant_conv(name, data, num_filter, kernel, stride, pad=(0, 0), no_bias=True, dilate=(1, 1), num_group=1,
               quant_mod="minmax", delay_quant=0, is_weight_perchannel=False):
    """quant_ops.py:81-107: <name>_weight -> <name>_weight_quant, data -> <name>_data_quant, Convolution <name>."""
    _need_mx()
    if is_weight_perchannel:
        assert quant_mod == "minmax", "currenet weight perchannel only support minmax node with weight"
    cin = get_sym_output_channel(name, data)
    weight = mx.sym.Variable(name=name + "_weight", dtype="float32",
                             shape=(num_filter, cin // num_group, kernel[0], kernel[1]))
    weight_q = _qnode(weight, name + "_weight_quant", True, quant_mod, delay_quant, is_weight_perchannel)
    data_q = _qnode(data, name + "_data_quant", False, quant_mod, delay_quant, False)
    return mx.symbol.Convolution(name=name, data=data_q, weight=weight_q, num_filter=num_filter, kernel=kernel,
                                 num_group=num_group, stride=stride, pad=pad, no_bias=no_bias, dilate=dilate)


def quant_fc(name, data, num_hidden, quant_mod="minmax", delay_quant=0, is_weight_perchannel=False):
    """quant_ops.py:109-121 (the FullyConnected node is literally named 'fc' there, :120)."""
    _need_mx()
    if is_weight_perchannel:
        assert quant_mod == "minmax", "currenet weight perchannel only support minmax node with weight"
    cin = get_sym_output_channel(name, data)
    weight = mx.sym.Variable(name=name + "_weight", shape=(num_hidden, cin), dtype="float32")
    weight_q = _qnode(weight, name + "_weight_quant", True, quant_mod, delay_quant, is_weight_perchannel)
    data_q = _qnode(data, name + "_data_quant", False, quant_mod, delay_quant, False)
    return mx.symbol.FullyConnected(data=data_q, num_hidden=num_hidden, name="fc", weight=weight_q)
