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

Two execution paths.  With ``quantized_op`` limited to the three layer types the model is rewritten in place, module by
module (works for any model).  As soon as ``quantized_op`` names one of ``Concat`` / ``Pooling`` / ``add_n`` /
``elemwise_add`` (graph_optimize.py:216-217,261-272) the model is traced with ``torch.fx`` and rewritten on its dataflow
graph, as the reference rewrites the symbol graph: every input of such an operator gets a data node, a tensor that feeds
several quantized operators is quantized once (keyed by its producer, :247-258,263-272), quantization nodes are named
after their producer, and a ``torch.fx.GraphModule`` is returned.  Called with an ``mx.sym`` symbol these functions are
the symbol-graph implementations of ``graph_optimize_sym``.
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


_GRAPH_OPS = ("Concat", "Pooling", "add_n", "elemwise_add")


def _is_symbol(obj):
    return hasattr(obj, "tojson") and hasattr(obj, "list_arguments")


class QuantizedWeightOp(nn.Module):
    """conv / FC / deconv whose WEIGHT is quantized here; its data input is quantized by a node of the dataflow graph
    (fx path), so that several consumers of one tensor share the node."""

    def __init__(self, name, inner, weight_node):
        super(QuantizedWeightOp, self).__init__()
        self.op_name = name
        self.inner = inner
        self.weight_quant = weight_node

    def forward(self, xq):
        wq = self.weight_quant(self.inner.weight)
        m = self.inner
        if isinstance(m, nn.Conv2d):
            return torch.nn.functional.conv2d(xq, wq, m.bias, m.stride, m.padding, m.dilation, m.groups)
        if isinstance(m, nn.ConvTranspose2d):
            return torch.nn.functional.conv_transpose2d(xq, wq, m.bias, m.stride, m.padding, m.output_padding, m.groups,
                                                        m.dilation)
        return torch.nn.functional.linear(xq, wq, m.bias)


def _fx_kind(node, modules):
    """reference operator name of an fx node (graph_optimize.py:216-217), or None."""
    import operator
    if node.op == "call_module":
        m = modules.get(node.target)
        if type(m) in _OP_KIND:
            return _OP_KIND[type(m)]
        if isinstance(m, (nn.MaxPool2d, nn.AvgPool2d, nn.AdaptiveAvgPool2d, nn.AdaptiveMaxPool2d)):
            return "Pooling"
        return None
    if node.op == "call_function":
        F = torch.nn.functional
        if node.target in (torch.cat, getattr(torch, "concat", torch.cat), getattr(torch, "concatenate", torch.cat)):
            return "Concat"
        if node.target in (F.max_pool2d, F.avg_pool2d, F.adaptive_avg_pool2d, F.adaptive_max_pool2d):
            return "Pooling"
        if node.target in (operator.add, operator.iadd, torch.add):
            tensors = [a for a in node.args if isinstance(a, torch.fx.Node)]
            return "elemwise_add" if len(tensors) == 2 else None
        if node.target is sum:
            return "add_n"
    if node.op == "call_method" and node.target in ("add", "add_"):
        return "elemwise_add" if len([a for a in node.args if isinstance(a, torch.fx.Node)]) == 2 else None
    return None


def _attach_fx(model, weight_setting, act_setting, quantized_op, skip_quantize_counts):
    import torch.fx as fx

    class Tracer(fx.Tracer):
        def is_leaf_module(self, m, qualname):
            return isinstance(m, (QuantNode, QuantizedOp, QuantizedWeightOp, Custom)) or super().is_leaf_module(m, qualname)

    graph = Tracer().trace(model)
    gm = fx.GraphModule(model, graph)
    modules = dict(gm.named_modules())
    visited = {k: 0 for k in ("Convolution", "FullyConnected", "Deconvolution") + _GRAPH_OPS}
    quantized = {}      # producer node name -> fx node of its quantization node (one per producer)

    def data_node(producer, before):
        if producer.name not in quantized:
            qname = producer.name + "_quant"
            gm.add_submodule(qname, create_quant_node(producer.name, act_setting))
            with gm.graph.inserting_before(before):
                quantized[producer.name] = gm.graph.call_module(qname, (producer,))
        return quantized[producer.name]

    for node in list(gm.graph.nodes):
        kind = _fx_kind(node, modules)
        if kind is None or kind not in quantized_op:
            continue
        visited[kind] += 1
        if skip_quantize_counts and kind in skip_quantize_counts and visited[kind] <= skip_quantize_counts[kind]:
            continue
        if kind in ("Convolution", "FullyConnected", "Deconvolution"):
            var = node.target.replace(".", "_")
            inner = modules[node.target]
            parent_name, _, child = node.target.rpartition(".")
            parent = modules[parent_name] if parent_name else gm
            setattr(parent, child, QuantizedWeightOp(var, inner, create_quant_node(var + "_weight", weight_setting)))
            producer = node.args[0]
            node.replace_input_with(producer, data_node(producer, node))
        else:
            tensors = []
            for a in node.args:
                if isinstance(a, fx.Node):
                    tensors.append(a)
                elif isinstance(a, (list, tuple)):
                    tensors.extend(x for x in a if isinstance(x, fx.Node))
            if kind == "Pooling":
                tensors = tensors[:1]
            for producer in tensors:
                if producer.op == "call_module" and isinstance(modules.get(producer.target), QuantNode):
                    continue
                node.replace_input_with(producer, data_node(producer, node))
    gm.graph.lint()
    gm.recompile()
    gm.quantized_op_counts = dict(visited)
    return gm


def attach_quantize_node(model, *args, **kwargs):
    """graph_optimize.py:199-292.  ``attach_quantize_node(model, weight_setting, act_setting, quantized_op=...,
    skip_quantize_counts=...)`` for a torch model; ``attach_quantize_node(symbol, out_shape_dict, weight_setting,
    act_setting, ...)`` for an ``mx.sym`` graph (the reference's signature)."""
    if _is_symbol(model):
        from . import graph_optimize_sym
        return graph_optimize_sym.attach_quantize_node(model, *args, **kwargs)
    return _attach_torch(model, *args, **kwargs)


def _attach_torch(model, weight_setting, act_setting,
                  quantized_op=("Convolution", "FullyConnected", "Deconvolution"), skip_quantize_counts=None):
    assert model is not None and weight_setting is not None and act_setting is not None
    if any(op in _GRAPH_OPS for op in quantized_op):
        return _attach_fx(model, weight_setting, act_setting, tuple(quantized_op), skip_quantize_counts)
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


class ChannelAffine(nn.Module):
    """What a folded BatchNorm becomes: ``x * gamma' + beta'`` with (1, C, 1, 1) parameters -- the reference's
    ``broadcast_add(BroadcastScale(x, gamma'), beta')`` (graph_optimize.py:91-93)."""

    def __init__(self, gamma, beta):
        super(ChannelAffine, self).__init__()
        self.gamma = nn.Parameter(gamma.reshape(1, -1, 1, 1).clone())
        self.beta = nn.Parameter(beta.reshape(1, -1, 1, 1).clone())

    def forward(self, x):
        return x * self.gamma + self.beta


def fix_bn(model):
    """graph_optimize.py:114-157: force use_global_stats on every BatchNorm (inference statistics while training)."""
    if _is_symbol(model):
        from . import graph_optimize_sym
        return graph_optimize_sym.fix_bn(model)
    for m in model.modules():
        if isinstance(m, nn.modules.batchnorm._BatchNorm):
            m.eval()
            m.train = lambda mode=True, _m=m: _m   # stays in eval mode
    return model


def _fold(bn):
    with torch.no_grad():
        inv = torch.sqrt(bn.running_var + bn.eps)
        gamma = bn.weight if bn.weight is not None else torch.ones_like(bn.running_var)   # affine=False / fix_gamma
        beta = bn.bias if bn.bias is not None else torch.zeros_like(bn.running_mean)
        new_beta = beta - gamma * bn.running_mean / inv          # beta first: it needs the unscaled gamma (:76-78)
        new_gamma = gamma / inv
        bn.running_mean.zero_()                                   # identity statistics: never folded twice (:86-87)
        bn.running_var.fill_(1.0)
    return ChannelAffine(new_gamma, new_beta)


def merge_bn(model, args=None, auxs=None, symbol_only=False):
    """graph_optimize.py:37-112.  A BatchNorm that uses its running statistics (``use_global_stats=True``: eval mode) and
    whose INPUT IS PRODUCED BY A CONVOLUTION is replaced by a per-channel scale and shift with
    gamma' = gamma / sqrt(eps + var), beta' = beta - gamma * mean / sqrt(eps + var); every other BatchNorm is left alone.
    The producer is found on the dataflow graph (``torch.fx``), not by attribute order; models fx cannot trace fall back
    to directly adjacent (conv, bn) pairs inside ``nn.Sequential`` containers.  Returns the model (a GraphModule on the
    fx path).  With an ``mx.sym`` symbol: the symbol-graph implementation, ``merge_bn(symbol, args, auxs, symbol_only)``."""
    if _is_symbol(model):
        from . import graph_optimize_sym
        return graph_optimize_sym.merge_bn(model, args, auxs, symbol_only)
    conv_like = (nn.Conv2d, QuantizedOp, QuantizedWeightOp)

    def is_conv(m):
        return isinstance(m, nn.Conv2d) or (isinstance(m, (QuantizedOp, QuantizedWeightOp)) and isinstance(m.inner, nn.Conv2d))

    def eligible(conv, bn):
        if bn.training or not bn.track_running_stats or bn.running_mean is None:
            return False                      # batch statistics: the reference folds only use_global_stats=True (:71)
        inner = conv.inner if isinstance(conv, (QuantizedOp, QuantizedWeightOp)) else conv
        assert inner.out_channels == bn.num_features, \
            "conv -> bn channel mismatch (%d vs %d)" % (inner.out_channels, bn.num_features)
        return True

    try:
        import torch.fx as fx

        class Tracer(fx.Tracer):
            def is_leaf_module(self, m, qualname):
                return isinstance(m, (QuantNode, QuantizedOp, QuantizedWeightOp, Custom)) or super().is_leaf_module(m, qualname)

        gm = fx.GraphModule(model, Tracer().trace(model))
    except Exception:
        gm = None
    if gm is not None:
        modules = dict(gm.named_modules())
        for node in gm.graph.nodes:
            if node.op != "call_module" or not isinstance(modules.get(node.target), nn.BatchNorm2d):
                continue
            src = node.args[0]
            if not (isinstance(src, fx.Node) and src.op == "call_module" and is_conv(modules.get(src.target))):
                continue
            bn = modules[node.target]
            if not eligible(modules[src.target], bn):
                continue
            parent_name, _, child = node.target.rpartition(".")
            setattr(modules[parent_name] if parent_name else gm, child, _fold(bn))
        gm.recompile()
        return gm
    for parent in model.modules():
        if not isinstance(parent, nn.Sequential):
            continue
        children = list(parent.named_children())
        for (n1, c1), (n2, c2) in zip(children, children[1:]):
            if is_conv(c1) and isinstance(c2, nn.BatchNorm2d) and eligible(c1, c2):
                setattr(parent, n2, _fold(c2))
    del conv_like
    return model
