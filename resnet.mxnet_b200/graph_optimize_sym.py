"""Symbol-graph flavour of ``core/graph_optimize.py`` (for MXNet users): the same four entry points on ``mx.sym``
graphs -- ``create_quant_node``, ``attach_quantize_node``, ``merge_bn``, ``fix_bn`` -- with the reference's naming and
traversal rules, instantiating THIS package's CustomOps.

All four walk the graph's JSON form in node order and rebuild it (the reference repeats that walker three times,
graph_optimize.py:44-112,121-157,210-290; here it is one function with a per-node hook).  Behaviour kept:

  create_quant_node   (:159-197)  node named like its input ``var.name``; learnable / aux variables ``<var>_minmax``,
                      ``<var>_alpha``, ``<var>_gamma``, ``<var>_pruning_point``, ``<var>_clipping_point`` with init values
                      0 / 1.0 / 8.0 (``setting.init_value`` overrides) and lr_mult 0.01, wd_mult 0 for the QIL points
  attach_quantize_node (:199-292) every variable must be in ``out_shape_dict`` and gets ``__shape__`` / ``__dtype__``;
                      ``skip_quantize_counts`` skips the first N ops of a type; Convolution / FullyConnected /
                      Deconvolution get a data node and a weight node; Concat / Pooling / add_n / elemwise_add get a data
                      node on EVERY input; a tensor feeding several quantized ops is quantized once (keyed by the
                      producer's name)
  merge_bn            (:37-112)   BatchNorm with use_global_stats=True fed by a Convolution becomes
                      broadcast_add(scale(data, gamma'), beta') with gamma' = gamma / sqrt(eps + var),
                      beta' = beta - gamma * mean / sqrt(eps + var), both reshaped to (1, C, 1, 1); moving stats are reset
                      to 0 / 1 so shared parameters are not folded twice
  fix_bn              (:114-157)  use_global_stats forced to True on every BatchNorm

Differences, all forced by what exists outside the fork: ``Quantization_int8`` builds the Python
``Quantization_int8_V2`` CustomOp instead of ``mx.sym.contrib.Quantization_int8`` (attributes the Python operator does
not have -- nbits, grad_mode, fix_act_scale -- are dropped); the ``*_CXX`` names build their Python twins with a warning;
``WNQ`` builds ``WNQ_PY`` (the reference accepts the name and then raises UnboundLocalError); the per-channel scale of
``merge_bn`` is ``contrib.BroadcastScale`` when the running MXNet has it and ``broadcast_mul`` otherwise.
"""
import json
import warnings

from ._mx import mx

FLOAT32_DTYPE = 0
_KNOWN = ("Quantization_int8", "QIL", "DoReFa_PY", "DoReFa_CXX", "PACT", "PACT_CXX", "WNQ", "GDRQ", "GDRQ_CXX")
_V2_ATTRS = ("quant_mode", "is_weight", "is_weight_perchannel", "delay_quant", "ema_decay")
_CONV_LIKE = ("Convolution", "FullyConnected", "Deconvolution")
_MULTI_IN = ("Concat", "Pooling", "add_n", "elemwise_add")


def _get(setting, key, default=None):
    if isinstance(setting, dict):
        return setting.get(key, default)
    return getattr(setting, key, default)


def get_constant(value):
    """'["constant", {"value": v}]' -- the serialised initializer the reference writes by hand (:32-34)."""
    return '["constant", {"value": ' + str(value) + '}]'


def _operator(op_name):
    if op_name.startswith("_contrib_"):
        return getattr(mx.sym.contrib, op_name[len("_contrib_"):])
    if op_name.startswith("_"):
        return getattr(mx.sym._internal, op_name)
    return getattr(mx.sym, op_name)


def _rebuild(symbol, on_variable=None, on_node=None):
    """Rebuild ``symbol`` node by node.  ``on_variable(name, attrs) -> attrs`` may edit a variable's double-underscore
    attributes; ``on_node(op, name, attrs, children, producers) -> symbol or None`` may replace a node (None: rebuild it
    unchanged).  ``producers[i]`` is the operator name that produced ``children[i]``."""
    assert symbol is not None
    graph = json.loads(symbol.tojson())
    built, built_op = {}, {}
    for nid, node in enumerate(graph["nodes"]):
        children = [built[e[0]][e[1]] for e in node["inputs"]]
        producers = [built_op[e[0]] for e in node["inputs"]]
        attrs = dict(node.get("attrs", {}))
        name, op = node["name"], node["op"]
        if op == "null":
            vattrs = {k: v for k, v in attrs.items() if k.startswith("__")}
            if on_variable is not None:
                vattrs = on_variable(name, vattrs)
            built[nid], built_op[nid] = mx.sym.var(name, **vattrs), "Variable"
            continue
        res = on_node(op, name, attrs, children, producers) if on_node is not None else None
        if res is None:
            res = _operator(op)(*children, **attrs, name=name)
        built[nid], built_op[nid] = res, (op[len("_contrib_"):] if op.startswith("_contrib_") else op)
    outs = [built[e[0]][e[1]] for e in graph["heads"]]
    return outs[0] if len(outs) == 1 else mx.sym.Group(outs)


# ---------------------------------------------------------------------------------------------------------------------
def create_quant_node(var, setting):
    """graph_optimize.py:159-197 on a symbol ``var``."""
    name = _get(setting, "quantize_op_name")
    attrs = dict(_get(setting, "attrs", {}) or {})
    init_value = _get(setting, "init_value", None)
    assert name in _KNOWN, "unknown quantize_op_name %r" % (name,)
    if name.endswith("_CXX"):
        twin = {"DoReFa_CXX": "DoReFa_PY", "PACT_CXX": "PACT", "GDRQ_CXX": "GDRQ"}[name]
        warnings.warn("%s is a C++ operator of the modified MXNet fork (README.md:7) whose source is not available; "
                      "building its Python twin %s (parity with the C++ operator is unpinned)" % (name, twin), stacklevel=2)
        name = twin
    vname = var.name
    if name == "Quantization_int8":
        minmax = mx.sym.var(name=vname + "_minmax", init=mx.init.Constant(init_value or 0))
        keep = {k: v for k, v in attrs.items() if k in _V2_ATTRS}
        return mx.sym.Custom(name=vname, data=var, minmax=minmax, op_type="Quantization_int8_V2", **keep)
    if name == "QIL":
        pruning = mx.sym.var(name=vname + "_pruning_point", init=get_constant(0), lr_mult=0.01, wd_mult=0)
        clipping = mx.sym.var(name=vname + "_clipping_point", init=get_constant(init_value or 1.0), lr_mult=0.01, wd_mult=0)
        gamma = mx.sym.var(name=vname + "_gamma", init=mx.init.Constant(1.0))
        return mx.sym.Custom(name=vname, data=var, pruning_point=pruning, clipping_point=clipping, gamma=gamma,
                             op_type="QIL_PY", **attrs)
    if name == "DoReFa_PY":
        return mx.sym.Custom(name=vname, data=var, op_type="DoReFa_PY", **attrs)
    if name == "PACT":
        gamma = mx.sym.var(name=vname + "_gamma", init=get_constant(init_value or 8.0))
        return mx.sym.Custom(name=vname, data=var, gamma=gamma, op_type="PACT_PY", **attrs)
    if name == "GDRQ":
        alpha = mx.sym.Variable(name=vname + "_alpha", init=mx.init.Constant(init_value or 1.0), dtype="float32")
        return mx.sym.Custom(name=vname, data=var, alpha=alpha, op_type="GDRQ_PY", **attrs)
    if name == "WNQ":
        keep = {k: v for k, v in attrs.items() if k in ("nbits", "is_perchannel")}
        return mx.sym.Custom(name=vname, data=var, op_type="WNQ_PY", **keep)
    raise AssertionError("unreachable: %r" % (name,))


def attach_quantize_node(symbol, out_shape_dict, weight_setting, act_setting,
                         quantized_op=("Convolution", "FullyConnected", "Deconvolution"), skip_quantize_counts=None):
    """graph_optimize.py:199-292."""
    assert weight_setting is not None and act_setting is not None
    visited = {k: 0 for k in _CONV_LIKE + _MULTI_IN}
    quantized = {}     # producer name -> its quantization node: a shared tensor is quantized once (:247-258)

    def shared(var, setting):
        if var.name not in quantized:
            quantized[var.name] = create_quant_node(var, setting)
        return quantized[var.name]

    def on_variable(name, attrs):
        assert name in out_shape_dict.keys(), "{} Variable is not in shape_dict".format(name)
        if "__shape__" not in attrs:
            attrs["__shape__"] = out_shape_dict[name]
            attrs["__dtype__"] = FLOAT32_DTYPE
        return attrs

    def on_node(op, name, attrs, children, producers):
        if op not in quantized_op:
            return None
        visited[op] += 1
        if skip_quantize_counts is not None and op in skip_quantize_counts and visited[op] <= skip_quantize_counts[op]:
            new_children = children                                                  # "skip idx" (:236-239)
        elif op in _CONV_LIKE:
            data, weight = children[0], children[1]
            bias = children[2] if len(children) > 2 else None
            new_children = [shared(data, act_setting), shared(weight, weight_setting), bias]
        elif op in _MULTI_IN:
            new_children = [shared(c, act_setting) for c in children]
        else:
            return None
        return _operator(op)(*new_children, **attrs, name=name)

    out = _rebuild(symbol, on_variable, on_node)
    attach_quantize_node.last_counts = dict(visited)
    return out


def fix_bn(symbol):
    """graph_optimize.py:114-157."""
    def on_node(op, name, attrs, children, producers):
        if op != "BatchNorm":
            return None
        if attrs.get("use_global_stats", "False") == "False":
            attrs["use_global_stats"] = "True"
        return mx.sym.BatchNorm(*children, **attrs, name=name)
    return _rebuild(symbol, None, on_node)


def _channel_scale(data, gamma):
    contrib = mx.sym.contrib
    try:
        op = contrib.BroadcastScale
    except AttributeError:
        return mx.sym.broadcast_mul(data, gamma)
    return op(data=data, scaler=gamma)


def merge_bn(symbol, args, auxs, symbol_only=False):
    """graph_optimize.py:37-112.  ``args`` / ``auxs``: name -> array maps (mx.nd, numpy or torch: only ``-``, ``*``,
    ``/``, ``** 0.5``, ``reshape`` and slice assignment are used); edited in place and returned like the reference does."""
    def on_node(op, name, attrs, children, producers):
        if op != "BatchNorm":
            return None
        _, gamma, beta, mmean, mvar = children
        g_name, b_name, m_name, v_name = gamma.name, beta.name, mmean.name, mvar.name
        assert "gamma" in g_name and "beta" in b_name and "moving_mean" in m_name and "moving_var" in v_name
        eps = float(attrs["eps"])
        if not (attrs.get("use_global_stats") == "True" and producers[0] == "Convolution"):
            return None
        if not symbol_only and m_name in auxs:
            inv_std = (auxs[v_name] + eps) ** 0.5
            # beta first: it needs the unscaled gamma
            new_beta = args[b_name] - args[g_name] * auxs[m_name] / inv_std
            new_gamma = args[g_name] / inv_std
            # stored broadcastable against NCHW, under the variables' own names and under <node>_gamma / <node>_beta
            # (the same key in the usual case; a copy when several BatchNorm nodes share parameters) (:79-90)
            args[g_name] = new_gamma.reshape((1, -1, 1, 1))
            args[b_name] = new_beta.reshape((1, -1, 1, 1))
            args[name + "_gamma"] = args[g_name]
            args[name + "_beta"] = args[b_name]
            auxs[m_name] = auxs[m_name].reshape((1, -1, 1, 1))
            auxs[v_name] = auxs[v_name].reshape((1, -1, 1, 1))
            auxs[m_name][:] = 0.0     # identity statistics: a shared BatchNorm must not be folded twice
            auxs[v_name][:] = 1.0
        shape_kw = {}
        if args is not None and (name + "_gamma") in args:
            shape_kw = {"shape": tuple(args[name + "_gamma"].shape)}
        new_gamma = mx.sym.var(name + "_gamma", **shape_kw)
        new_beta = mx.sym.var(name + "_beta", **shape_kw)
        return mx.sym.broadcast_add(_channel_scale(children[0], new_gamma), new_beta)
    out = _rebuild(symbol, None, on_node)
    return out, args, auxs
