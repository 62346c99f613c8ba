"""Drop-in for ``symbol/fold_bn_v1_gdrq.py``: op_type ``GDRQ_Fold_BN``.

forward (fold_bn_v1_gdrq.py:32-120), three library calls and one convolution:
    data   : 2*mean|x| reduction with the first-batch-init / EMA update of aux[0] fused in, then one sweep that
             clips with the *batch* threshold and scales with the *EMA* one (:53-68)
    weight : gamma/sqrt(var+eps) folded in on the fly, 2*mean|w'| per tensor or per out-channel, clip, QDQ, and
             the folded bias beta - mean*gamma/sqrt(var+eps), in two launches with no per-channel host sync
             (the reference loops over channels with two asnumpy() each, :84-85)
    conv   : library convolution (cuDNN through torch, or mx.nd.Convolution under MXNet) + bias (:99-120)
backward (:122-129): gradient only to ``bn_output``; the six other inputs get zeros.
"""
import os

from . import _kernels as K
from . import _lib
from .operator import CustomOp, CustomOpProp, py_bool, py_literal, register

try:
    import torch
    import torch.nn.functional as F
except Exception:  # pragma: no cover
    torch = None

ALLOW_TF32_CONV = os.environ.get("B2Q_CONV_TF32", "0") == "1"


class GDRQ_Fold_BN(CustomOp):
    def __init__(self, quant_mode, is_weight_perchannel, delay_quant, ema_decay, name, num_filter, num_group,
                 kernel, stride, pad, dilate, no_bias, eps, momentum, fix_gamma, quantize_flag):
        self.quant_mode = quant_mode
        self.is_weight_perchannel = is_weight_perchannel
        self.delay_quant = delay_quant
        self.ema_decay = ema_decay
        self.QUANT_LEVEL = 127
        self.init = True
        self.sync = None        # dist.ThresholdSync: max over ranks of mean|data| before the EMA
        self.peer = None        # dist.PeerThresholdExchange: the same exchange fused into the two data kernels
        self.name = name
        self.num_filter = num_filter
        self.num_group = num_group
        self.kernel = kernel
        self.stride = stride
        self.pad = pad
        self.dilate = dilate
        self.no_bias = no_bias
        assert self.no_bias == True, "fold bn don't support bias mode in conv or deconv"  # noqa: E712  (:24)
        self.eps = eps
        self.momentum = momentum
        self.fix_gamma = fix_gamma
        self.quantize_flag = quantize_flag
        # kept for inspection by tests / callers
        self.data_q = None
        self.weight_q = None
        self.bias = None

    def _conv(self, data, weight, bias):
        if torch is not None and isinstance(data, torch.Tensor):
            # true float32 convolution like the reference's cuDNN v5/v6 call (README.md:8); TF32 is opt-in
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=ALLOW_TF32_CONV):
                return F.conv2d(data, weight, bias, stride=tuple(self.stride), padding=tuple(self.pad),
                                dilation=tuple(self.dilate), groups=int(self.num_group))
        import mxnet as mx  # pragma: no cover
        conv = mx.nd.Convolution(name=self.name, data=data, weight=weight, num_filter=self.num_filter,
                                 kernel=self.kernel, num_group=self.num_group, stride=self.stride, pad=self.pad,
                                 dilate=self.dilate, no_bias=True)
        return conv + bias.reshape((1, -1, 1, 1))

    @staticmethod
    def _empty_like(a, shape=None):
        if torch is not None and isinstance(a, torch.Tensor):
            return torch.empty(tuple(shape) if shape is not None else a.shape, dtype=a.dtype, device=a.device)
        import mxnet as mx  # pragma: no cover
        return mx.nd.empty(tuple(shape) if shape is not None else a.shape, ctx=a.context, dtype=a.dtype)

    def forward(self, is_train, req, in_data, out_data, aux):
        assert len(in_data) == 7, \
            "fold bn require seven inputs: data, weight, bn_output, bn_gamma, bn_beta, bn_mean, bn_var"  # :33
        data, weight, bn_output, bn_gamma, bn_beta, bn_mean, bn_var = in_data
        if is_train and self.delay_quant > 0:          # :43-47
            self.assign(out_data[0], req[0], bn_output)
            self.delay_quant -= 1
            return
        if self.quantize_flag:
            if not is_train:
                # the reference only binds `thresholds` when training (:55-58) and then reads it (:67)
                raise NameError("name 'thresholds' is not defined")
            data_q = self._empty_like(data)
            if self.peer is not None:
                self.peer.quantize_mean(_lib.UPD_TWICE_STORE if self.init else _lib.UPD_TWICE_EMA, data, data_q, aux[0],
                                        self.ema_decay, 1 - self.ema_decay, 127.0)
            elif self.sync is not None:
                # data parallel: mean|x| -> allreduce(max) -> t = 2*mean, EMA (:58-64) -> clip by t, scale by aux (:65-68)
                if getattr(self, "_stat", None) is None:
                    self._stat = self._empty_like(aux[0], (1,))
                    self._clip = self._empty_like(aux[0], (1,))
                K.meanabs(data, self._stat)
                self.sync(self._stat)
                K.threshold_update(_lib.UPD_TWICE_STORE if self.init else _lib.UPD_TWICE_EMA, self._stat, aux[0],
                                   self.ema_decay, 1 - self.ema_decay, clip_out=self._clip)
                K.qdq(data, data_q, aux[0], 127.0, _lib.CLIP_SYM, "write", clip_thr=self._clip)
            else:
                K.foldbn_data_fwd(data, data_q, aux[0], self.init, self.ema_decay)   # :53-68
            self.init = False
        else:
            data_q = data
        pre = getattr(self, "_prefolded", None)
        if pre is not None:
            # the caller produced bn_mean / bn_var WITH the fused statistics + fold kernel
            # (b2q_bnstat_foldbn_weight_fwd_f32, harness.FoldBNConv2d): the weight path is already done for this call
            weight_q, bias = pre
            self._prefolded = None
        else:
            weight_q = self._empty_like(weight)
            bias = self._empty_like(bn_beta, (self.num_filter,))
            K.foldbn_weight_fwd(weight, weight_q, bias, aux[1], bn_gamma, bn_beta, bn_mean, bn_var, self.eps,
                                self.is_weight_perchannel, self.quantize_flag, is_train)   # :70-96, :113
        self.data_q, self.weight_q, self.bias = data_q, weight_q, bias
        self.assign(out_data[0], req[0], self._conv(data_q, weight_q, bias))   # :99-120

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        for g in in_grad:                               # :124-125
            K.zero_(g)
        self.assign(in_grad[2], req[2], out_grad[0])    # :129


@register("GDRQ_Fold_BN")
class GDRQ_FoldBNProp(CustomOpProp):
    """fold_bn_v1_gdrq.py:131-191."""

    def __init__(self, quant_mode, is_weight_perchannel=False, delay_quant=0, ema_decay=0.99, name="fold_bn",
                 num_filter=None, num_group=None, kernel=(3, 3), stride=(1, 1), pad=(0, 0), dilate=(1, 1),
                 no_bias=True, eps=1e-5, momentum=0.9, fix_gamma=False, quantize_flag="True"):
        self.quant_mode = str(quant_mode)
        self.delay_quant = int(delay_quant)
        self.ema_decay = float(ema_decay)
        self.is_weight_perchannel = py_bool(is_weight_perchannel)
        self.name = str(name)
        self.num_filter = int(num_filter)
        self.num_group = int(num_group)
        self.kernel = tuple(py_literal(kernel))
        self.stride = tuple(py_literal(stride))
        self.pad = tuple(py_literal(pad))
        self.dilate = tuple(py_literal(dilate))
        self.no_bias = py_bool(no_bias)
        self.eps = float(eps)
        self.momentum = float(momentum)
        self.fix_gamma = py_bool(fix_gamma)
        self.quantize_flag = py_bool(quantize_flag)
        super(GDRQ_FoldBNProp, self).__init__(True)

    def list_arguments(self):
        return ["data", "weight", "bn_output", "bn_gamma", "bn_beta", "bn_mean", "bn_var"]

    def list_outputs(self):
        return ["output"]

    def list_auxiliary_states(self):
        return ["data_minmax", "weight_minmax"]

    def infer_shape(self, in_shape):
        shape = in_shape[0]
        aux_shape = [[1], [in_shape[1][0]] if self.is_weight_perchannel else [1]]
        oshape = [None] * len(shape)
        oshape[0] = shape[0]
        oshape[1] = self.num_filter
        for ax in (0, 1):   # :179-182
            oshape[2 + ax] = int((shape[2 + ax] + 2 * self.pad[ax] -
                                  (self.dilate[ax] * (self.kernel[ax] - 1) + 1)) / self.stride[ax] + 1)
        return in_shape, [oshape], aux_shape

    def infer_type(self, in_type):
        return in_type, [in_type[0]] * len(self.list_outputs()), [in_type[0]] * len(self.list_auxiliary_states())

    def create_operator(self, ctx, shapes, dtypes):
        return GDRQ_Fold_BN(self.quant_mode, self.is_weight_perchannel, self.delay_quant, self.ema_decay, self.name,
                            self.num_filter, self.num_group, self.kernel, self.stride, self.pad, self.dilate,
                            self.no_bias, self.eps, self.momentum, self.fix_gamma, self.quantize_flag)


def GDRQ_fold_bn(name, data, quant_mod="minmax", is_weight_perchannel=False, delay_quant=0, ema_decay=0.99,
                 num_filter=None, kernel=None, stride=None, pad=(0, 0), no_bias=True, dilate=(1, 1), num_group=1,
                 w_lr_mult=None, w_wd_mult=None, w_init=None, eps=1e-5, momentum=0.9, fix_gamma=False,
                 use_global_stats=False, gamma_lr_mult=None, gamma_wd_mult=None, gamma_init=None,
                 beta_lr_mult=None, beta_wd_mult=None, beta_init=None, quantize_flag=True, dict_shapes=None):
    """Symbol builder of fold_bn_v1_gdrq.py:237-288 (needs MXNet): conv + BatchNorm_v1(output_mean_var) feeding
    the fold-BN custom op.  Quirks kept: ``delay_quant=0`` is hard-wired into the op (:279) and gamma/beta are
    declared with the *input* channel count (:263-266).  Under torch use ``b200quant.harness.FoldBNConv2d``."""
    from ._mx import mx
    mx.module()
    if is_weight_perchannel:
        assert quant_mod == "minmax", "currenet weight perchannel only support minmax node with weight"
    assert dict_shapes is not None, "please setting dict_shapes for infer shape"
    args = data.list_arguments()
    _, out_shapes, _ = data.infer_shape(**{k: v for k, v in dict_shapes.items() if k in args})
    cin = out_shapes[0][1]
    weight = mx.sym.Variable(name=name + "_conv2d_weight", shape=(num_filter, cin // num_group, kernel[0], kernel[1]),
                             dtype="float32", lr_mult=w_lr_mult, wd_mult=w_wd_mult, init=w_init)
    gamma = mx.symbol.Variable(name + "_batchnorm_gamma", shape=(cin,), dtype="float32", lr_mult=gamma_lr_mult,
                               wd_mult=gamma_wd_mult, init=gamma_init)
    beta = mx.symbol.Variable(name + "_batchnorm_beta", shape=(cin,), dtype="float32", lr_mult=beta_lr_mult,
                              wd_mult=beta_wd_mult, init=beta_init)
    conv = mx.sym.Convolution(name=name + "_conv2d", data=data, weight=weight, num_filter=num_filter, kernel=kernel,
                              num_group=num_group, stride=stride, pad=pad, no_bias=no_bias, dilate=dilate)
    bn_out, bn_mean, bn_var = mx.sym.BatchNorm_v1(name=name + "_batchnorm", data=conv, gamma=gamma, beta=beta, eps=eps,
                                                  momentum=momentum, fix_gamma=fix_gamma, output_mean_var=True,
                                                  use_global_stats=use_global_stats)
    return mx.sym.Custom(name=name + "_fold_bn", data=data, weight=weight, bn_output=bn_out, bn_gamma=gamma,
                         bn_beta=beta, bn_mean=bn_mean, bn_var=bn_var, quant_mode=quant_mod,
                         is_weight_perchannel=is_weight_perchannel, delay_quant=0, ema_decay=ema_decay,
                         num_filter=num_filter, num_group=num_group, kernel=kernel, stride=stride, pad=pad,
                         dilate=dilate, no_bias=no_bias, eps=eps, momentum=momentum, fix_gamma=fix_gamma,
                         quantize_flag=quantize_flag, op_type="GDRQ_Fold_BN")
