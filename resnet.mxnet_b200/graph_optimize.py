"""Drop-in (torch-module flavour) for the reference's ``core/graph_optimize.py``: the *live* way the reference puts the
quantization operators around every conv / FC (``train.py:112-120``).

``attach_quantize_node(model, weight_setting, act_setting, quantized_op, skip_quantize_counts)`` walks a torch model
in definition order and wraps every ``nn.Conv2d`` ("Convolution"), ``nn.Linear`` ("FullyConnected") and
``nn.ConvTranspose2d`` ("Deconvolution") with a data-quant node and a weight-quant node created by
``create_quant_node`` from the same ``{"quantize_op_name", "init_value", "attrs"}`` settings the reference reads from
``config.quantize_setting`` (``config/edict_config.py:158-187``, schema ``config/quant_attrs.py``).  Behaviour kept
from ``graph_optimize.py:199-292``:
  * ``skip_quantize_counts`` skips the first N ops of a type (``:236-239``);
  * a tensor that feeds several quantized ops is quantized ONCE (``:247-258``): the first consumer's data node owns the
    threshold, later consumers reuse its output;
  * aux / parameter names: ``<var>_minmax``, ``<var>_alpha``, ``<var>_gamma``, ``<var>_pruning_point``,
    ``<var>_clipping_point`` with init values 0 / 1.0 / 8.0 and ``lr_mult=0.01, wd_mult=0`` for QIL points (``:166-194``).
``quantize_op_name`` mapping (``:165-195``): Quantization_int8 -> the Python Quantization_int8_V2 (the reference maps it
to the fork's C++ contrib op, whose source is absent: SURVEY.md F3), QIL -> QIL_PY, DoReFa_PY, PACT -> PACT_PY,
GDRQ -> GDRQ_PY, WNQ -> WNQ_PY (the reference accepts the name and then fails with UnboundLocalError, ``:162-163``);
the ``*_CXX`` names (``mx.sym.contrib.{DoReFa,PACT,GDRQ}`` of the fork, source absent) run the Python twin's
arithmetic on the same kernels, with a one-time warning that their parity is unpinned.
``fix_bn`` / ``merge_bn`` are the inference helpers of ``graph_optimize.py:37-157``.
"""
import warnings

import torch
import torch.nn as nn

from .harness import Custom

_OP_KIND = {nn.Conv2d: "Convolution", nn.Linear: "FullyConnected", nn.ConvTranspose2d: "Deconvolution"}
_KNOWN = ("Quantization_int8", "QIL", "DoReFa_PY", "DoReFa_CXX", "PACT", "PACT_CXX", "WNQ", "GDRQ", "GDRQ_CXX")


def _get(setting, key, default=None):
    if isinstance(setting, dict):
        return setting.get(key, default)
    return getattr(setting, key, default)


class QuantNode(nn.Module):
    """One quantization node made by ``create_quant_node``: the CustomOp plus the learnable scalars some ops take as
    *arguments* (PACT gamma, QIL points) with the reference's names, init values and lr/wd multipliers."""

    def __init__(self, var_name, op_type, attrs, aux_init=None, params=None):
        super(QuantNode, self).__init__()
        self.var_name = var_name
        self.node = Custom(op_type, aux_init=aux_init if aux_init is not None else 1.0, **attrs)
        self.param_names = []
        for pname, value, lr_mult, wd_mult, trainable in (params or []):
            p = nn.Parameter(torch.full((1,), float(value)), requires_grad=trainable)
            p.lr_mult, p.wd_mult = lr_mult, wd_mult
            self.register_parameter(pname, p)
            self.param_names.append(pname)

    def forward(self, x):
        extra = [getattr(self, n) for n in self.param_names]
        return self.node(x, *extra)

    def mx_names(self):
        args = {self.var_name + "_" + n: getattr(self, n) for n in self.param_names}
        aux = {self.var_name + "_" + n: getattr(self.node, n) for n in self.node.aux_names
               if getattr(self.node, n, None) is not None}
        return args, aux


def create_quant_node(var_name, setting):
    """graph_optimize.py:159-197."""
    name = _get(setting, "quantize_op_name")
    attrs = dict(_get(setting, "attrs", {}) or {})
    init_value = _get(setting, "init_value", None)
    assert name in _KNOWN, "unknown quantize_op_name %r" % (name,)
    if name.endswith("_CXX"):
        warnings.warn("%s is a C++ operator of the modified MXNet fork (README.md:7) whose source is not available; "
                      "using the arithmetic of its Python twin %s (parity with the C++ operator is unpinned)"
                      % (name, {"DoReFa_CXX": "DoReFa_PY", "PACT_CXX": "PACT_PY", "GDRQ_CXX": "GDRQ_PY"}[name]),
                      stacklevel=2)
        name = {"DoReFa_CXX": "DoReFa_PY", "PACT_CXX": "PACT", "GDRQ_CXX": "GDRQ"}[name]
    if name == "Quantization_int8":
        # the C++ op's extra attributes (nbits, grad_mode, fix_act_scale) have no counterpart in the Python op
        keep = {k: v for k, v in attrs.items()
                if k in ("quant_mode", "is_weight", "is_weight_perchannel", "delay_quant", "ema_decay")}
        return QuantNode(var_name, "Quantization_int8_V2", keep, aux_init=init_value or 0)
    if name == "QIL":
        return QuantNode(var_name, "QIL_PY", attrs, params=[("pruning_point", 0.0, 0.01, 0.0, True),
                                                            ("clipping_point", init_value or 1.0, 0.01, 0.0, True),
                                                            ("gamma", 1.0, 1.0, 1.0, False)])
    if name == "DoReFa_PY":
        return QuantNode(var_name, "DoReFa_PY", attrs)
    if name == "PACT":
        return QuantNode(var_name, "PACT_PY", attrs, params=[("gamma", init_value or 8.0, 1.0, 1.0, True)])
    if name == "GDRQ":
        return QuantNode(var_name, "GDRQ_PY", attrs, aux_init=init_value or 1.0)
    if name == "WNQ":
        keep = {k: v for k, v in attrs.items() if k in ("nbits", "is_perchannel")}
        return QuantNode(var_name, "WNQ_PY", keep)
    raise AssertionError("unreachable: %r" % (name,))


class _Share(object):
    """Per-forward cache so that a tensor feeding several quantized ops is quantized once (graph_optimize.py:247-258)."""

    def __init__(self):
        self.cache = {}

    def clear(self, *_):
        self.cache.clear()


class QuantizedOp(nn.Module):
    """A conv / FC / deconv with its attached quantization nodes."""

    def __init__(self, name, inner, data_node, weight_node, share):
        super(QuantizedOp, self).__init__()
        self.op_name = name
        self.inner = inner
        self.data_quant = data_node
        self.weight_quant = weight_node
        self._share = [share]   # list: keep it out of the module tree

    def forward(self, x):
        share = self._share[0].cache
        key = id(x)
        hit = share.get(key)
        if hit is not None and hit[0] is x:
            xq = hit[1]
        else:
            xq = self.data_quant(x)
            share[key] = (x, xq)
        wq = self.weight_quant(self.inner.weight)
        m = self.inner
        if isinstance(m, nn.Conv2d):
            return torch.nn.functional.conv2d(xq, wq, m.bias, m.stride, m.padding, m.dilation, m.groups)
        if isinstance(m, nn.ConvTranspose2d):
            return torch.nn.functional.conv_transpose2d(xq, wq, m.bias, m.stride, m.padding, m.output_padding, m.groups,
                                                        m.dilation)
        return torch.nn.functional.linear(xq, wq, m.bias)


def attach_quantize_node(model, weight_setting, act_setting,
                         quantized_op=("Convolution", "FullyConnected", "Deconvolution"), skip_quantize_counts=None):
    """graph_optimize.py:199-292 for a torch model (modified in place and returned)."""
    assert model is not None and weight_setting is not None and act_setting is not None
    visited = {"Convolution": 0, "FullyConnected": 0, "Deconvolution": 0}
    share = _Share()
    model.register_forward_pre_hook(share.clear)
    targets = []
    for parent_name, parent in model.named_modules():
        for child_name, child in parent.named_children():
            kind = _OP_KIND.get(type(child))
            if kind is None or kind not in quantized_op:
                continue
            targets.append((parent, child_name, (parent_name + "." if parent_name else "") + child_name, child, kind))
    for parent, child_name, full_name, child, kind in targets:
        visited[kind] += 1
        if skip_quantize_counts and kind in skip_quantize_counts and visited[kind] <= skip_quantize_counts[kind]:
            continue   # "skip idx:{} {} on {}" (:236-239)
        var = full_name.replace(".", "_")
        wrapped = QuantizedOp(var, child, create_quant_node(var + "_data", act_setting),
                              create_quant_node(var + "_weight", weight_setting), share)
        setattr(parent, child_name, wrapped)
    model.quantized_op_counts = dict(visited)
    return model


def export_quant_params(model):
    """(arg_params, aux_params) of the attached nodes in the rewriter's naming (``<var>_minmax`` ...)."""
    args, aux = {}, {}
    for m in model.modules():
        if isinstance(m, QuantNode):
            a, x = m.mx_names()
            args.update({k: v.detach() for k, v in a.items()})
            aux.update({k: v.detach() for k, v in x.items()})
    return args, aux


def fix_bn(model):
    """graph_optimize.py:114-157: force use_global_stats on every BatchNorm (inference statistics while training)."""
    for m in model.modules():
        if isinstance(m, nn.modules.batchnorm._BatchNorm):
            m.eval()
            m.train = lambda mode=True, _m=m: _m   # stays in eval mode
    return model


def merge_bn(model):
    """graph_optimize.py:37-112: fold every conv -> BatchNorm pair for inference (scale into the weight, shift into the
    bias), in place.  Only directly adjacent pairs inside nn.Sequential containers or attribute pairs named
    (<x>, <x>_bn / bn<k>) are folded."""
    def fold(conv, bn):
        with torch.no_grad():
            f = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            conv.weight.mul_(f.reshape(-1, 1, 1, 1))
            bias = conv.bias if conv.bias is not None else torch.zeros_like(bn.running_mean)
            conv.bias = nn.Parameter(bn.bias + (bias - bn.running_mean) * f)

    for parent in model.modules():
        children = list(parent.named_children())
        for (n1, c1), (n2, c2) in zip(children, children[1:]):
            if isinstance(c1, nn.Conv2d) and isinstance(c2, nn.BatchNorm2d):
                fold(c1, c2)
                setattr(parent, n2, nn.Identity())
    return model
