"""torch driver for the CustomOp classes (the only runnable host in an image without MXNet).

``Custom`` plays the role of ``mx.sym.Custom(..., op_type=...)`` plus MXNet's executor for one node: it builds
the operator through its registered Prop from *string* attributes, owns the auxiliary states as buffers named
like the reference's (``minmax``, ``alpha``, ``data_minmax`` ...), and calls ``forward`` / ``backward`` with
the MXNet protocol (lists of arrays, ``req`` strings, ``is_train``).  torch only supplies device memory, the
stream and autograd bookkeeping; all arithmetic is in libb2q.so.
"""
import torch
import torch.nn as nn

from .operator import get_prop


class _CustomFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, node, *inputs):
        in_data = [t.detach().contiguous() for t in inputs]
        out = node._alloc_out(in_data)
        node.op.forward(bool(node.training), ["write"], in_data, [out], node.aux_list())
        ctx.node = node
        # save_for_backward, not attributes: an output kept as ctx.out would form a reference cycle through the
        # autograd node and keep every activation of the step alive
        ctx.save_for_backward(*in_data, out)
        ctx.needs = [bool(t.requires_grad) for t in inputs]
        return out

    @staticmethod
    def backward(ctx, grad_out):
        node = ctx.node
        saved = ctx.saved_tensors
        in_data, out = list(saved[:-1]), saved[-1]
        if node.alias_ste and node.is_identity_backward():
            return (None, grad_out) + (None,) * (len(in_data) - 1)
        grads = [torch.empty_like(t) for t in in_data]
        node.op.backward(["write"] * len(in_data), [grad_out.contiguous()], in_data, [out], grads,
                         node.aux_list())
        return (None,) + tuple(g if need else None for g, need in zip(grads, ctx.needs))


class Custom(nn.Module):
    """One quantization node.  ``Custom("Quantization_int8_V2", is_weight=True, quant_mode="minmax")(w)``."""

    def __init__(self, op_type, aux_init=1.0, alias_ste=False, **attrs):
        super(Custom, self).__init__()
        self.op_type = op_type
        self.attrs = {k: str(v) for k, v in attrs.items()}   # MXNet hands attributes over as strings
        self.prop = get_prop(op_type)(**self.attrs)
        self.op = self.prop.create_operator(None, None, None)
        self.aux_names = list(self.prop.list_auxiliary_states())
        self.aux_init = aux_init
        self.alias_ste = alias_ste
        self._aux_ready = False

    # -- auxiliary states ------------------------------------------------------------------------------
    def _ensure_aux(self, in_data):
        if self._aux_ready:
            return
        _, _, aux_shapes = self.prop.infer_shape([list(t.shape) for t in in_data])
        init = self.aux_init if isinstance(self.aux_init, (list, tuple)) else [self.aux_init] * len(self.aux_names)
        for name, shape, v in zip(self.aux_names, aux_shapes, init):
            if getattr(self, name, None) is None:
                self.register_buffer(name, torch.full(tuple(shape), float(v), dtype=torch.float32,
                                                      device=in_data[0].device))
        self._aux_ready = True

    def aux_list(self):
        return [getattr(self, n) for n in self.aux_names]

    def _alloc_out(self, in_data):
        self._ensure_aux(in_data)
        _, out_shapes, _ = self.prop.infer_shape([list(t.shape) for t in in_data])
        return torch.empty(tuple(out_shapes[0]), dtype=torch.float32, device=in_data[0].device)

    def is_identity_backward(self):
        op = self.op
        if self.op_type in ("Quantization_int8_V2", "QUANT_STE_PY"):
            return True
        if self.op_type in ("ClipGrad_Quantization_int8", "GDRQ_PY"):
            return bool(op.is_weight)
        return False

    # -- op-instance state the reference forgets to checkpoint ---------------------
    def get_extra_state(self):
        return {k: getattr(self.op, k) for k in ("delay_quant", "init") if hasattr(self.op, k)}

    def set_extra_state(self, state):
        for k, v in (state or {}).items():
            setattr(self.op, k, v)

    def forward(self, *inputs):
        return _CustomFn.apply(self, *inputs)

    def extra_repr(self):
        return "op_type=%s, %s" % (self.op_type, ", ".join("%s=%s" % kv for kv in sorted(self.attrs.items())))


# ----------------------------------------------------------------------------------------------------------
# Layer wrappers with the reference's naming contract
# ----------------------------------------------------------------------------------------------------------
_STYLES = {
    # symbol/quant_ops.py:87-94: nodes <name>_weight_quant / <name>_data_quant, op Quantization_int8_V2
    "quant_ops": ("Quantization_int8_V2", "_weight_quant", "_data_quant"),
    # symbol/int8_api.py:29-36: nodes <name>_weight / <name>_data, op ClipGrad_Quantization_int8
    "int8_api": ("ClipGrad_Quantization_int8", "_weight", "_data"),
}


class _QuantLayer(nn.Module):
    def _make_nodes(self, name, style, quant_mod, delay_quant, is_weight_perchannel, ema_decay):
        if is_weight_perchannel:
            assert quant_mod == "minmax", "currenet weight perchannel only support minmax node with weight"
        op_type, wsuf, dsuf = _STYLES[style]
        self.layer_name = name
        self.weight_node_name = name + wsuf
        self.data_node_name = name + dsuf
        self.weight_quant = Custom(op_type, quant_mode=quant_mod, is_weight=True,
                                   is_weight_perchannel=is_weight_perchannel, ema_decay=ema_decay,
                                   delay_quant=delay_quant)
        self.data_quant = Custom(op_type, quant_mode=quant_mod, is_weight=False, is_weight_perchannel=False,
                                 ema_decay=ema_decay, delay_quant=delay_quant)

    def mx_names(self):
        """MXNet checkpoint names of this layer's parameters / aux states (arg:<..>, aux:<..>)."""
        args = {self.layer_name + "_weight": self.weight}
        if getattr(self, "bias", None) is not None:
            args[self.layer_name + "_bias"] = self.bias
        aux = {}
        for node, mod in ((self.weight_node_name, self.weight_quant), (self.data_node_name, self.data_quant)):
            for aname in mod.aux_names:
                if getattr(mod, aname, None) is not None:
                    aux[node + "_" + aname] = getattr(mod, aname)
        return args, aux


class QuantConv2d(_QuantLayer):
    """``quant_conv`` (symbol/quant_ops.py:81-107) / ``clipgrad_quant_conv`` (symbol/int8_api.py:19-50) as a torch
    module: weight Variable ``<name>_weight`` -> weight-quant node, input -> data-quant node, Convolution."""

    def __init__(self, name, in_channels, num_filter, kernel, stride=(1, 1), pad=(0, 0), no_bias=True, dilate=(1, 1),
                 num_group=1, quant_mod="minmax", delay_quant=0, is_weight_perchannel=False, ema_decay=0.99,
                 style="quant_ops"):
        super(QuantConv2d, self).__init__()
        self.stride, self.pad, self.dilate, self.num_group = tuple(stride), tuple(pad), tuple(dilate), int(num_group)
        self.weight = nn.Parameter(torch.empty(num_filter, in_channels // num_group, kernel[0], kernel[1]))
        nn.init.kaiming_normal_(self.weight, mode="fan_in", nonlinearity="relu")   # Xavier(gaussian, in, 2), train.py:221
        self.bias = None if no_bias else nn.Parameter(torch.zeros(num_filter))
        self._make_nodes(name, style, quant_mod, delay_quant, is_weight_perchannel, ema_decay)

    def forward(self, x):
        return torch.nn.functional.conv2d(self.data_quant(x), self.weight_quant(self.weight), self.bias, self.stride,
                                          self.pad, self.dilate, self.num_group)


class QuantLinear(_QuantLayer):
    """``quant_fc`` (symbol/quant_ops.py:109-121) / ``clipgrad_quant_fc`` (symbol/int8_api.py:52-71)."""

    def __init__(self, name, in_features, num_hidden, no_bias=False, quant_mod="minmax", delay_quant=0,
                 is_weight_perchannel=False, ema_decay=0.99, style="quant_ops"):
        super(QuantLinear, self).__init__()
        self.weight = nn.Parameter(torch.empty(num_hidden, in_features))
        nn.init.kaiming_normal_(self.weight, mode="fan_in", nonlinearity="relu")
        self.bias = None if no_bias else nn.Parameter(torch.zeros(num_hidden))
        self._make_nodes(name, style, quant_mod, delay_quant, is_weight_perchannel, ema_decay)

    def forward(self, x):
        return torch.nn.functional.linear(self.data_quant(x.flatten(1)), self.weight_quant(self.weight), self.bias)


class QuantDeconv2d(_QuantLayer):
    """``clipgrad_quant_deconv`` (symbol/int8_api.py:73-94): weight Variable ``<name>_weight`` of shape
    (in_channels, num_filter, kh, kw), nodes ``<name>_weight`` / ``<name>_data``, then Deconvolution."""

    def __init__(self, name, in_channels, kernel, stride, pad, num_filter, no_bias=True, quant_mod="minmax",
                 delay_quant=0, is_weight_perchannel=False, ema_decay=0.99, style="int8_api"):
        super(QuantDeconv2d, self).__init__()
        self.stride, self.pad = tuple(stride), tuple(pad)
        self.weight = nn.Parameter(torch.empty(in_channels, num_filter, kernel[0], kernel[1]))
        nn.init.kaiming_normal_(self.weight, mode="fan_in", nonlinearity="relu")
        self.bias = None if no_bias else nn.Parameter(torch.zeros(num_filter))
        self._make_nodes(name, style, quant_mod, delay_quant, is_weight_perchannel, ema_decay)

    def forward(self, x):
        return torch.nn.functional.conv_transpose2d(self.data_quant(x), self.weight_quant(self.weight), self.bias,
                                                    self.stride, self.pad)


class _DataNodes(nn.Module):
    """Activation-only wrappers of symbol/int8_api.py:96-117: ClipGrad data nodes in front of an add / concat, or alone."""

    def _node(self, node_name, quant_mode, delay_quant, ema_decay):
        q = Custom("ClipGrad_Quantization_int8", quant_mode=quant_mode, is_weight=False, is_weight_perchannel=False,
                   ema_decay=ema_decay, delay_quant=delay_quant)
        q.node_name = node_name
        return q

    def mx_names(self):
        aux = {}
        for mod in self.children():
            if isinstance(mod, Custom):
                for aname in mod.aux_names:
                    if getattr(mod, aname, None) is not None:
                        aux[mod.node_name + "_" + aname] = getattr(mod, aname)
        return {}, aux


class QuantData(_DataNodes):
    """``clipgrad_quant_data`` (int8_api.py:96-99): node ``<name>_data``."""

    def __init__(self, name, quant_mode="minmax", delay_quant=0, ema_decay=0.99):
        super(QuantData, self).__init__()
        self.data_quant = self._node(name + "_data", quant_mode, delay_quant, ema_decay)

    def forward(self, x):
        return self.data_quant(x)


class QuantAdd(_DataNodes):
    """``clipgrad_quant_add`` (int8_api.py:101-108): nodes ``<name>add_lhs_data`` / ``<name>add_rhs_data`` (no separator,
    as in the reference), sum ``<name>_plus``."""

    def __init__(self, name, quant_mode="minmax", delay_quant=0, ema_decay=0.99):
        super(QuantAdd, self).__init__()
        self.out_name = name + "_plus"
        self.lhs_quant = self._node(name + "add_lhs_data", quant_mode, delay_quant, ema_decay)
        self.rhs_quant = self._node(name + "add_rhs_data", quant_mode, delay_quant, ema_decay)

    def forward(self, lhs, rhs):
        return self.lhs_quant(lhs) + self.rhs_quant(rhs)


class QuantConcat(_DataNodes):
    """``clipgrad_quant_concat`` (int8_api.py:110-117): nodes ``<name>concat_{i}_data``, then concat along ``dim``."""

    def __init__(self, name, num_inputs, dim=1, quant_mode="minmax", delay_quant=0, ema_decay=0.99):
        super(QuantConcat, self).__init__()
        self.dim = dim
        self.out_name = name
        self.quants = nn.ModuleList([self._node(name + "concat_{}_data".format(i), quant_mode, delay_quant, ema_decay)
                                     for i in range(num_inputs)])

    def mx_names(self):
        aux = {}
        for mod in self.quants:
            for aname in mod.aux_names:
                if getattr(mod, aname, None) is not None:
                    aux[mod.node_name + "_" + aname] = getattr(mod, aname)
        return {}, aux

    def forward(self, inputs):
        assert isinstance(inputs, (list, tuple)), "the input fo quantize concat must be a list"
        assert len(inputs) == len(self.quants)
        return torch.cat([q(x) for q, x in zip(self.quants, inputs)], dim=self.dim)


class FoldBNConv2d(nn.Module):
    """``GDRQ_fold_bn`` (symbol/fold_bn_v1_gdrq.py:237-288) as a torch module: convolution -> BatchNorm with batch
    statistics (``BatchNorm_v1(output_mean_var=True)``) -> the ``GDRQ_Fold_BN`` operator fed (data, weight, bn_output,
    gamma, beta, mean, var).  Names: weight ``<name>_conv2d_weight``, ``<name>_batchnorm_gamma`` / ``_beta`` (+ moving
    statistics), node ``<name>_fold_bn`` with aux ``data_minmax`` / ``weight_minmax``.  The gradient reaches the
    parameters through ``bn_output`` only, as in the reference (:122-129).  gamma / beta have ``num_filter`` elements
    (the reference declares them with the INPUT channel count, :263-266, which only binds when the two agree).
    ``fused=True`` (default): batch statistics, fold, per-channel weight quantisation and bias come from ONE launch
    (``b2q_bnstat_foldbn_weight_fwd_f32``); ``fused=False`` computes the statistics with the
    stand-alone kernel and lets the operator run its own weight path -- same bits."""

    def __init__(self, name, in_channels, num_filter, kernel, stride=(1, 1), pad=(0, 0), dilate=(1, 1), num_group=1,
                 quant_mod="minmax", is_weight_perchannel=False, ema_decay=0.99, eps=1e-5, momentum=0.9,
                 fix_gamma=False, use_global_stats=False, quantize_flag=True, fused=True):
        super(FoldBNConv2d, self).__init__()
        self.layer_name = name
        self.stride, self.pad, self.dilate, self.num_group = tuple(stride), tuple(pad), tuple(dilate), int(num_group)
        self.eps, self.momentum, self.use_global_stats, self.fused = float(eps), float(momentum), use_global_stats, fused
        self.per_channel, self.quantize_flag = bool(is_weight_perchannel), bool(quantize_flag)
        self.weight = nn.Parameter(torch.empty(num_filter, in_channels // num_group, kernel[0], kernel[1]))
        nn.init.kaiming_normal_(self.weight, mode="fan_in", nonlinearity="relu")
        self.gamma = nn.Parameter(torch.ones(num_filter), requires_grad=not fix_gamma)
        self.beta = nn.Parameter(torch.zeros(num_filter))
        self.register_buffer("moving_mean", torch.zeros(num_filter))
        self.register_buffer("moving_var", torch.ones(num_filter))
        self.fold_bn = Custom("GDRQ_Fold_BN", quant_mode=quant_mod, is_weight_perchannel=is_weight_perchannel,
                              delay_quant=0, ema_decay=ema_decay, name=name, num_filter=num_filter, num_group=num_group,
                              kernel=tuple(kernel), stride=tuple(stride), pad=tuple(pad), dilate=tuple(dilate),
                              no_bias=True, eps=eps, momentum=momentum, fix_gamma=fix_gamma, quantize_flag=quantize_flag)

    def mx_names(self):
        n = self.layer_name
        args = {n + "_conv2d_weight": self.weight, n + "_batchnorm_gamma": self.gamma, n + "_batchnorm_beta": self.beta}
        aux = {n + "_batchnorm_moving_mean": self.moving_mean, n + "_batchnorm_moving_var": self.moving_var}
        for aname in self.fold_bn.aux_names:
            if getattr(self.fold_bn, aname, None) is not None:
                aux[n + "_fold_bn_" + aname] = getattr(self.fold_bn, aname)
        return args, aux

    def forward(self, x):
        from . import _kernels as K
        conv = torch.nn.functional.conv2d(x, self.weight, None, self.stride, self.pad, self.dilate, self.num_group)
        c = conv.shape[1]
        if self.training and not self.use_global_stats:
            mean, var = torch.empty(c, device=x.device), torch.empty(c, device=x.device)
            cd = conv.detach().contiguous()
            if self.fused and self.quantize_flag and self.per_channel:
                node = self.fold_bn
                probe = [x.detach(), self.weight.detach(), cd, self.gamma.detach(), self.beta.detach(), mean, var]
                node._ensure_aux(probe)
                w_q, bias = torch.empty_like(self.weight), torch.empty(c, device=x.device)
                K.bnstat_foldbn_weight_fwd(cd, mean, var, self.weight.detach(), w_q, bias, node.weight_minmax,
                                           self.gamma.detach(), self.beta.detach(), self.eps, True, True, True)
                node.op._prefolded = (w_q, bias)
            else:
                K.bn_batch_stats(cd, mean, var)
            with torch.no_grad():   # [upstream batch_norm_v1-inl.h] moving = moving * momentum + batch * (1 - momentum)
                self.moving_mean.mul_(self.momentum).add_(mean, alpha=1 - self.momentum)
                self.moving_var.mul_(self.momentum).add_(var, alpha=1 - self.momentum)
        else:
            mean, var = self.moving_mean, self.moving_var
        shape = (1, -1, 1, 1)
        bn_out = (conv - mean.view(shape)) / torch.sqrt(var.view(shape) + self.eps) * self.gamma.view(shape) \
            + self.beta.view(shape)
        return self.fold_bn(x, self.weight, bn_out, self.gamma, self.beta, mean, var)


def export_mx_params(model):
    """(arg_params, aux_params) name -> tensor maps in the reference's checkpoint naming (train.py:218,224-227),
    plus the per-op Python state the reference forgets to save (delay_quant countdown, first-batch init flag)."""
    arg_params, aux_params, op_state = {}, {}, {}
    for mod in model.modules():
        if isinstance(mod, _QuantLayer):
            a, x = mod.mx_names()
            arg_params.update({k: v.detach() for k, v in a.items()})
            aux_params.update({k: v.detach() for k, v in x.items()})
            op_state[mod.weight_node_name] = mod.weight_quant.get_extra_state()
            op_state[mod.data_node_name] = mod.data_quant.get_extra_state()
        elif isinstance(mod, FoldBNConv2d):
            a, x = mod.mx_names()
            arg_params.update({k: v.detach() for k, v in a.items()})
            aux_params.update({k: v.detach() for k, v in x.items()})
            op_state[mod.layer_name + "_fold_bn"] = mod.fold_bn.get_extra_state()
        elif isinstance(mod, _DataNodes):
            _, x = mod.mx_names()
            aux_params.update({k: v.detach() for k, v in x.items()})
            for q in mod.modules():
                if isinstance(q, Custom):
                    op_state[q.node_name] = q.get_extra_state()
    return arg_params, aux_params, op_state


class SimpleCifarNet(nn.Module):
    """symbol/simple.py:10-18 with every conv / FC routed through quant_conv / quant_fc (BASELINE.json config 1)."""

    def __init__(self, num_classes=10, style="quant_ops", **quant_kw):
        super(SimpleCifarNet, self).__init__()
        self.stage1_conv = QuantConv2d("stage1_conv", 3, 8, (3, 3), (2, 2), (1, 1), True, style=style, **quant_kw)
        self.stage1_bn = nn.BatchNorm2d(8, eps=1e-3, momentum=0.1)   # mx BatchNorm defaults: eps 1e-3, momentum 0.9
        self.stage2_conv = QuantConv2d("stage2_conv", 8, 8, (3, 3), (2, 2), (1, 1), True, style=style, **quant_kw)
        self.stage2_bn = nn.BatchNorm2d(8, eps=1e-3, momentum=0.1)
        self.fc1 = QuantLinear("fc1", 8, num_classes, style=style, **quant_kw)

    def forward(self, x):
        x = torch.relu(self.stage1_bn(self.stage1_conv(x)))
        x = torch.relu(self.stage2_bn(self.stage2_conv(x)))
        x = x.mean(dim=(2, 3))                       # global average pool (simple.py:14)
        return self.fc1(x)


class _ResidualUnitInt8(nn.Module):
    """residual_unit_int8 (symbol/resnet_int8.py:12-67): pre-activation unit, every conv through quant_conv; the
    shortcut conv of a dim-changing unit quantizes act1 again with its own node (:38)."""

    def __init__(self, name, channel, num_filter, stride, dim_match, bottle_neck, style, bn_mom=0.9, eps=1e-5, **qkw):
        super(_ResidualUnitInt8, self).__init__()
        self.dim_match, self.bottle_neck = dim_match, bottle_neck
        mom = 1.0 - bn_mom
        if bottle_neck:
            mid = int(num_filter * 0.25)
            self.bn1 = nn.BatchNorm2d(channel, eps=eps, momentum=mom)
            self.conv1 = QuantConv2d(name + "_conv1", channel, mid, (1, 1), (1, 1), (0, 0), True, style=style, **qkw)
            self.bn2 = nn.BatchNorm2d(mid, eps=eps, momentum=mom)
            self.conv2 = QuantConv2d(name + "_conv2", mid, mid, (3, 3), stride, (1, 1), True, style=style, **qkw)
            self.bn3 = nn.BatchNorm2d(mid, eps=eps, momentum=mom)
            self.conv3 = QuantConv2d(name + "_conv3", mid, num_filter, (1, 1), (1, 1), (0, 0), True, style=style, **qkw)
        else:
            self.bn1 = nn.BatchNorm2d(channel, eps=eps, momentum=mom)
            self.conv1 = QuantConv2d(name + "_conv1", channel, num_filter, (3, 3), stride, (1, 1), True, style=style, **qkw)
            self.bn2 = nn.BatchNorm2d(num_filter, eps=eps, momentum=mom)
            self.conv2 = QuantConv2d(name + "_conv2", num_filter, num_filter, (3, 3), (1, 1), (1, 1), True, style=style, **qkw)
        self.sc = None if dim_match else QuantConv2d(name + "_sc", channel, num_filter, (1, 1), stride, (0, 0), True,
                                                     style=style, **qkw)

    def forward(self, x):
        act1 = torch.relu(self.bn1(x))
        if self.bottle_neck:
            out = self.conv1(act1)
            out = self.conv2(torch.relu(self.bn2(out)))
            out = self.conv3(torch.relu(self.bn3(out)))
        else:
            out = self.conv1(act1)
            out = self.conv2(torch.relu(self.bn2(out)))
        return out + (x if self.sc is None else self.sc(act1))


class ResNetInt8(nn.Module):
    """resnet_int8 (symbol/resnet_int8.py:69-131) as a torch benchmark harness: the convolutions / BatchNorm / pooling
    are library code; every conv and the FC go through the quantization operators of this package."""

    def __init__(self, units=(3, 4, 6, 3), filter_list=(64, 256, 512, 1024, 2048), num_classes=1000, bottle_neck=True,
                 dataset_type="imagenet", style="quant_ops", bn_mom=0.9, **qkw):
        super(ResNetInt8, self).__init__()
        self.dataset_type = dataset_type
        self.bn_data = nn.BatchNorm2d(3, eps=2e-5, momentum=1.0 - bn_mom, affine=False)    # fix_gamma=True (:86)
        if dataset_type == "imagenet":
            self.conv0 = QuantConv2d("conv0", 3, filter_list[0], (7, 7), (2, 2), (3, 3), True, style=style, **qkw)
            self.bn0 = nn.BatchNorm2d(filter_list[0], eps=1e-5, momentum=1.0 - bn_mom)
        else:
            self.conv0 = QuantConv2d("conv0", 3, filter_list[0], (3, 3), (1, 1), (1, 1), True, style=style, **qkw)
        blocks = []
        for i, n_units in enumerate(units):
            stride = (1, 1) if i == 0 else (2, 2)
            blocks.append(_ResidualUnitInt8("stage%d_unit1" % (i + 1), filter_list[i], filter_list[i + 1], stride, False,
                                            bottle_neck, style, bn_mom, **qkw))
            for j in range(n_units - 1):
                blocks.append(_ResidualUnitInt8("stage%d_unit%d" % (i + 1, j + 2), filter_list[i + 1], filter_list[i + 1],
                                                (1, 1), True, bottle_neck, style, bn_mom, **qkw))
        self.blocks = nn.Sequential(*blocks)
        self.bn1 = nn.BatchNorm2d(filter_list[-1], eps=1e-5, momentum=1.0 - bn_mom)
        self.fc1 = QuantLinear("fc1", filter_list[-1], num_classes, style=style, **qkw)

    def forward(self, x):
        x = self.conv0(self.bn_data(x))
        if self.dataset_type == "imagenet":
            x = torch.nn.functional.max_pool2d(torch.relu(self.bn0(x)), 3, 2, 1)
        x = self.blocks(x)
        x = torch.relu(self.bn1(x)).mean(dim=(2, 3))
        return self.fc1(x)


def quant_nodes(model):
    """All Custom quantization nodes of a model, in definition order."""
    return [m for m in model.modules() if isinstance(m, Custom)]
