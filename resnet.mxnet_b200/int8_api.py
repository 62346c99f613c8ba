"""Drop-in for ``symbol/int8_api.py``: symbol-level wrappers that wire a weight Variable, a weight-quant node, a
data-quant node and the conv / FC / deconv / add / concat together.

Naming contract kept (int8_api.py:29-36): parameter ``<name>_weight``, quant nodes ``<name>_weight`` and
``<name>_data`` (so their aux states are ``<name>_weight_minmax`` / ``<name>_data_minmax``), layer ``<name>``.
``clipgrad_quant_*`` use the ``ClipGrad_Quantization_int8`` CustomOp of this package.  ``quant_*_cxx`` target the
fork's C++ ``contrib.Quantization_int8`` (int8_api.py:133-209), whose source is not in the reference tree
(parity unpinned): they are provided for API completeness and only work on that MXNet fork.
These are graph builders and need MXNet's symbolic API; under torch use ``b200quant.harness``.
"""
from .clip_grad_quantization_int8 import *  # noqa: F401,F403  (registers ClipGrad_Quantization_int8)
from .quant_ops import _need_mx

from ._mx import mx

_OP = "ClipGrad_Quantization_int8"


def get_sym_output_channel(name, sym, dict_shapes=None):
    _need_mx()
    assert dict_shapes is not None, "please setting dict_shapes for infer shape"
    args = sym.list_arguments()
    _, out_shapes, _ = sym.infer_shape(**{k: v for k, v in dict_shapes.items() if k in args})
    assert len(out_shapes) == 1, "the output of sym is not equal to 1"
    return out_shapes[0][1]


def _check(init, is_weight_perchannel, quant_mode):
    _need_mx()
    if init is not None:
        assert isinstance(init, mx.init.Initializer)
    if is_weight_perchannel:
        assert quant_mode == "minmax", "currenet weight perchannel only support minmax node with weight"


def _py_node(data, name, is_weight, quant_mode, delay_quant, ema_decay, perchannel=False):
    return mx.sym.Custom(data=data, name=name, quant_mode=quant_mode, is_weight=is_weight,
                         is_weight_perchannel=perchannel, ema_decay=ema_decay, delay_quant=delay_quant, op_type=_OP)


def _cxx_node(data, name, is_weight, quant_mode, delay_quant, ema_decay, grad_mode, workspace, perchannel=False):
    if not hasattr(mx.sym.contrib, "Quantization_int8"):
        raise RuntimeError("mx.sym.contrib.Quantization_int8 is only available in the modified MXNet fork "
                           "(README.md:7); use the clipgrad_quant_* / quant_ops wrappers instead")
    return mx.sym.contrib.Quantization_int8(data=data, name=name, quant_mode=quant_mode, is_weight=is_weight,
                                            is_weight_perchannel=perchannel, ema_decay=ema_decay,
                                            delay_quant=delay_quant, grad_mode=grad_mode, workspace=workspace)


def _weight_var(name, shape, weight, lr_mult, wd_mult, init):
    if isinstance(weight, mx.sym.Symbol):
        return weight
    return mx.sym.Variable(name=name + "_weight", shape=shape, dtype="float32", lr_mult=lr_mult, wd_mult=wd_mult,
                           init=init)


def _conv_like(node, name, data, num_filter, kernel, stride, pad, no_bias, dilate, num_group, lr_mult, wd_mult, init,
               weight, bias, quant_mode, delay_quant, is_weight_perchannel, ema_decay, dict_shapes):
    _check(init, is_weight_perchannel, quant_mode)
    cin = get_sym_output_channel(name, data, dict_shapes=dict_shapes)
    weight = _weight_var(name, (num_filter, cin // num_group, kernel[0], kernel[1]), weight, lr_mult, wd_mult, init)
    weight_q = node(weight, name + "_weight", True, quant_mode, delay_quant, ema_decay, is_weight_perchannel)
    data_q = node(data, name + "_data", False, quant_mode, delay_quant, ema_decay)
    return mx.symbol.Convolution(name=name, data=data_q, num_filter=num_filter, kernel=kernel, num_group=num_group,
                                 stride=stride, pad=pad, no_bias=no_bias, dilate=dilate, weight=weight_q, bias=bias)


def _fc_like(node, name, data, num_hidden, flatten, no_bias, lr_mult, wd_mult, init, weight, bias, quant_mode,
             delay_quant, is_weight_perchannel, ema_decay, dict_shapes):
    _check(init, is_weight_perchannel, quant_mode)
    cin = get_sym_output_channel(name, data, dict_shapes=dict_shapes)
    weight = _weight_var(name, (num_hidden, cin), weight, lr_mult, wd_mult, init)
    weight_q = node(weight, name + "_weight", True, quant_mode, delay_quant, ema_decay, is_weight_perchannel)
    data_q = node(data, name + "_data", False, quant_mode, delay_quant, ema_decay)
    return mx.symbol.FullyConnected(data=data_q, num_hidden=num_hidden, name=name, weight=weight_q, flatten=flatten,
                                    no_bias=no_bias, bias=bias)


def _deconv_like(node, name, data, kernel, stride, pad, num_filter, no_bias, cudnn_tune, lr_mult, wd_mult, init, weight,
                 bias, quant_mode, delay_quant, is_weight_perchannel, ema_decay, dict_shapes):
    _check(init, is_weight_perchannel, quant_mode)
    cin = get_sym_output_channel(name, data, dict_shapes=dict_shapes)
    weight = _weight_var(name, (cin, num_filter, kernel[0], kernel[1]), weight, lr_mult, wd_mult, init)
    weight_q = node(weight, name + "_weight", True, quant_mode, delay_quant, ema_decay, is_weight_perchannel)
    data_q = node(data, name + "_data", False, quant_mode, delay_quant, ema_decay)
    return mx.symbol.Deconvolution(name=name, data=data_q, kernel=kernel, stride=stride, pad=pad, no_bias=no_bias,
                                   num_filter=num_filter, cudnn_tune=cudnn_tune, weight=weight_q, bias=bias)


# ---- ClipGrad (Python CustomOp) family: int8_api.py:19-117 --------------------------------------------------
def clipgrad_quant_conv(name, data, num_filter, kernel, stride, pad=(0, 0), no_bias=False, dilate=(1, 1), num_group=1,
                        lr_mult=None, wd_mult=None, init=None, weight=None, bias=None, quant_mode="minmax",
                        delay_quant=0, is_weight_perchannel=False, ema_decay=0.99, dict_shapes=None):
    return _conv_like(_py_node, name, data, num_filter, kernel, stride, pad, no_bias, dilate, num_group, lr_mult,
                      wd_mult, init, weight, bias, quant_mode, delay_quant, is_weight_perchannel, ema_decay, dict_shapes)


def clipgrad_quant_fc(name, data, num_hidden, flatten=True, no_bias=False, lr_mult=None, wd_mult=None, init=None,
                      weight=None, bias=None, quant_mode="minmax", delay_quant=0, is_weight_perchannel=False,
                      ema_decay=0.99, dict_shapes=None):
    return _fc_like(_py_node, name, data, num_hidden, flatten, no_bias, lr_mult, wd_mult, init, weight, bias,
                    quant_mode, delay_quant, is_weight_perchannel, ema_decay, dict_shapes)


def clipgrad_quant_deconv(name, data, kernel, stride, pad, num_filter, no_bias=True, cudnn_tune="fastest",
                          lr_mult=None, wd_mult=None, init=None, weight=None, bias=None, quant_mode="minmax",
                          delay_quant=0, is_weight_perchannel=False, ema_decay=0.99, dict_shapes=None):
    return _deconv_like(_py_node, name, data, kernel, stride, pad, num_filter, no_bias, cudnn_tune, lr_mult, wd_mult,
                        init, weight, bias, quant_mode, delay_quant, is_weight_perchannel, ema_decay, dict_shapes)


def clipgrad_quant_data(name, data, quant_mode="minmax", delay_quant=0, ema_decay=0.99):
    _need_mx()
    return mx.sym.Custom(data=data, name=name + "_data", quant_mode=quant_mode, is_weight=False, ema_decay=ema_decay,
                         delay_quant=delay_quant, op_type=_OP)


def clipgrad_quant_add(name, lhs_data, rhs_data, quant_mode="minmax", delay_quant=0, ema_decay=0.99):
    _need_mx()
    lhs = _py_node(lhs_data, name + "add_lhs_data", False, quant_mode, delay_quant, ema_decay)
    rhs = _py_node(rhs_data, name + "add_rhs_data", False, quant_mode, delay_quant, ema_decay)
    return mx.symbol.ElementWiseSum(lhs, rhs, name=name + "_plus")


def clipgrad_quant_concat(name, inputs, dim=1, quant_mode="minmax", delay_quant=0, ema_decay=0.99):
    _need_mx()
    assert isinstance(inputs, list), "the input fo quantize concat must be a list"
    qs = [_py_node(x, name + "concat_{}_data".format(i), False, quant_mode, delay_quant, ema_decay)
          for i, x in enumerate(inputs)]
    return mx.symbol.concat(*qs, dim=dim, name=name)


# ---- fork-only C++ family: int8_api.py:120-209 ----------------------------------------------------------------
def _cxx(grad_mode, workspace):
    def node(data, name, is_weight, quant_mode, delay_quant, ema_decay, perchannel=False):
        return _cxx_node(data, name, is_weight, quant_mode, delay_quant, ema_decay, grad_mode, workspace, perchannel)
    return node


def quant_conv_cxx(name, data, num_filter, kernel, stride, pad=(0, 0), no_bias=False, dilate=(1, 1), num_group=1,
                   lr_mult=None, wd_mult=None, init=None, weight=None, bias=None, quant_mode="minmax", delay_quant=0,
                   is_weight_perchannel=False, ema_decay=0.99, grad_mode="ste", workspace=512, dict_shapes=None):
    return _conv_like(_cxx(grad_mode, workspace), name, data, num_filter, kernel, stride, pad, no_bias, dilate,
                      num_group, lr_mult, wd_mult, init, weight, bias, quant_mode, delay_quant, is_weight_perchannel,
                      ema_decay, dict_shapes)


def quant_fc_cxx(name, data, num_hidden, flatten=True, no_bias=False, lr_mult=None, wd_mult=None, init=None,
                 weight=None, bias=None, quant_mode="minmax", delay_quant=0, is_weight_perchannel=False,
                 ema_decay=0.99, grad_mode="ste", workspace=512, dict_shapes=None):
    return _fc_like(_cxx(grad_mode, workspace), name, data, num_hidden, flatten, no_bias, lr_mult, wd_mult, init,
                    weight, bias, quant_mode, delay_quant, is_weight_perchannel, ema_decay, dict_shapes)


def quant_deconv_cxx(name, data, kernel, stride, pad, num_filter, no_bias=True, cudnn_tune="fastest", lr_mult=None,
                     wd_mult=None, init=None, weight=None, bias=None, quant_mode="minmax", delay_quant=0,
                     is_weight_perchannel=False, ema_decay=0.99, grad_mode="ste", workspace=512, dict_shapes=None):
    return _deconv_like(_cxx(grad_mode, workspace), name, data, kernel, stride, pad, num_filter, no_bias, cudnn_tune,
                        lr_mult, wd_mult, init, weight, bias, quant_mode, delay_quant, is_weight_perchannel, ema_decay,
                        dict_shapes)


def quant_add_cxx(name, lhs_data, rhs_data, quant_mode="minmax", delay_quant=0, ema_decay=0.99, grad_mode="ste",
                  workspace=512):
    _need_mx()
    node = _cxx(grad_mode, workspace)
    lhs = node(lhs_data, name + "add_lhs_data", False, quant_mode, delay_quant, ema_decay)
    rhs = node(rhs_data, name + "add_rhs_data", False, quant_mode, delay_quant, ema_decay)
    return mx.symbol.ElementWiseSum(lhs, rhs, name=name + "_plus")


def quant_concat_cxx(name, inputs, dim=1, quant_mode="minmax", delay_quant=0, ema_decay=0.99, grad_mode="ste",
                     workspace=512):
    _need_mx()
    assert isinstance(inputs, list), "the input fo quantize concat must be a list"
    node = _cxx(grad_mode, workspace)
    qs = [node(x, name + "concat_{}_data".format(i), False, quant_mode, delay_quant, ema_decay)
          for i, x in enumerate(inputs)]
    return mx.symbol.concat(*qs, dim=dim, name=name)
