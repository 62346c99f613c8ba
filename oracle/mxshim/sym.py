"""TEST INFRASTRUCTURE ONLY -- a minimal ``mx.sym`` for the shim: a symbolic graph recorder with MXNet's naming, JSON
layout and shape inference for the handful of operators the reference's graph builders touch.

Purpose: the reference's symbol-level code -- ``quant_conv`` / ``quant_fc`` (symbol/quant_ops.py:81-121), the
``clipgrad_quant_*`` / ``quant_*_cxx`` wrappers (symbol/int8_api.py:19-209), ``GDRQ_fold_bn``
(symbol/fold_bn_v1_gdrq.py:237-288) and ``create_quant_node`` / ``attach_quantize_node`` / ``merge_bn`` / ``fix_bn``
(core/graph_optimize.py:37-292) -- runs unmodified over it, next to this repository's builders, and
tests/test_symbol_builders.py diffs the two graphs (node names, operators, attributes, argument and aux names).

What it models [upstream python/mxnet/symbol/symbol.py, nnvm JSON]:
  * ``tojson()``: {"nodes": [{"op", "name", "attrs", "inputs": [[node_id, output_index, 0], ...]}], "arg_nodes",
    "heads"}; variables have op "null"; contrib operators are stored as ``_contrib_<Name>``; ``ElementWiseSum`` is an
    alias of ``add_n``, ``concat`` of ``Concat``; attribute values are strings;
  * missing parameter inputs are created as variables named ``<node>_<input>`` (``conv0_weight``, ``bn0_gamma``,
    ``bn0_moving_mean`` -- the last two of BatchNorm are auxiliary states); a Custom node takes its argument / aux names
    from the registered CustomOpProp;
  * ``list_arguments`` / ``list_auxiliary_states`` in depth-first input order; ``infer_shape(**known)``.
"""
import json

_REGISTRIES = []      # dicts op_type -> CustomOpProp class, searched in order (set_registries)
_COUNTERS = {}


def set_registries(*regs):
    del _REGISTRIES[:]
    _REGISTRIES.extend(regs)


def reset_names():
    _COUNTERS.clear()


def _auto_name(hint):
    i = _COUNTERS.get(hint, 0)
    _COUNTERS[hint] = i + 1
    return "%s%d" % (hint, i)


def _attr_str(v):
    if isinstance(v, str):
        return v
    if isinstance(v, (list, tuple)):
        return "(" + ", ".join(str(x) for x in v) + ("," if len(v) == 1 else "") + ")"
    if hasattr(v, "dumps"):
        return v.dumps()
    return str(v)


def _tuple(s, default=None):
    if s is None:
        return default
    if isinstance(s, (tuple, list)):
        return tuple(int(x) for x in s)
    s = s.strip().strip("()[]")
    return tuple(int(float(x)) for x in s.split(",") if x.strip())


def _bool(s, default=False):
    if s is None:
        return default
    return str(s).strip() in ("True", "true", "1")


class _Node(object):
    def __init__(self, op, name, attrs, inputs, num_outputs=1, is_aux=False):
        self.op, self.name, self.attrs, self.inputs = op, name, attrs, inputs   # inputs: [(node, out_index)]
        self.num_outputs = num_outputs
        self.is_aux = is_aux          # variable that is an auxiliary state of its consumer


class Symbol(object):
    def __init__(self, heads):
        self._heads = heads           # [(node, out_index)]

    # -- identity -----------------------------------------------------------------------------------------------
    @property
    def name(self):
        if len(self._heads) != 1:
            return None
        return self._heads[0][0].name

    def __getitem__(self, i):
        if isinstance(i, int):
            if len(self._heads) == 1 and self._heads[0][0].num_outputs > 1:
                return Symbol([(self._heads[0][0], i)])
            return Symbol([self._heads[i]])
        raise TypeError("only integer output indices are modelled")

    def __iter__(self):
        node = self._heads[0][0]
        if len(self._heads) == 1 and node.num_outputs > 1:
            return iter([Symbol([(node, i)]) for i in range(node.num_outputs)])
        return iter([Symbol([h]) for h in self._heads])

    def __len__(self):
        node = self._heads[0][0]
        return node.num_outputs if len(self._heads) == 1 else len(self._heads)

    # -- arithmetic [upstream: elemwise_add / _plus_scalar ...] ---------------------------------------------------
    def __add__(self, other):
        if isinstance(other, Symbol):
            return _make("elemwise_add", [self, other], {}, None, hint="_plus")
        return _make("_plus_scalar", [self], {"scalar": other}, None, hint="_plusscalar")

    __radd__ = __add__

    def __mul__(self, other):
        if isinstance(other, Symbol):
            return _make("elemwise_mul", [self, other], {}, None, hint="_mul")
        return _make("_mul_scalar", [self], {"scalar": other}, None, hint="_mulscalar")

    __rmul__ = __mul__

    # -- traversal ----------------------------------------------------------------------------------------------
    def _topo(self):
        order, seen = [], set()

        def visit(node):
            if id(node) in seen:
                return
            seen.add(id(node))
            for src, _ in node.inputs:
                visit(src)
            order.append(node)
        for node, _ in self._heads:
            visit(node)
        return order

    def list_arguments(self):
        return [n.name for n in self._topo() if n.op == "null" and not n.is_aux]

    def list_auxiliary_states(self):
        return [n.name for n in self._topo() if n.op == "null" and n.is_aux]

    def list_outputs(self):
        out = []
        for node, i in self._heads:
            names = _output_names(node)
            out.append(node.name if node.op == "null" else "%s_%s" % (node.name, names[i]))
        return out

    def get_internals(self):
        heads = []
        for n in self._topo():
            heads.extend((n, i) for i in range(n.num_outputs))
        return Symbol(heads)

    def attr_dict(self):
        return {n.name: dict(n.attrs) for n in self._topo() if n.attrs}

    def tojson(self):
        order = self._topo()
        index = {id(n): i for i, n in enumerate(order)}
        nodes = []
        for n in order:
            d = {"op": n.op, "name": n.name, "inputs": [[index[id(s)], i, 0] for s, i in n.inputs]}
            if n.attrs:
                d["attrs"] = {k: _attr_str(v) for k, v in n.attrs.items()}
            nodes.append(d)
        return json.dumps({"nodes": nodes, "arg_nodes": [i for i, n in enumerate(order) if n.op == "null"],
                           "heads": [[index[id(n)], i, 0] for n, i in self._heads],
                           "attrs": {"mxnet_version": ["int", 10500]}}, indent=2)

    # -- shapes -------------------------------------------------------------------------------------------------
    def infer_shape(self, **known):
        shapes = {}     # id(node) -> [shape per output]

        def var_shape(node):
            if node.name in known:
                return tuple(known[node.name])
            s = node.attrs.get("__shape__")
            return _tuple(s) if s is not None else None

        for n in self._topo():
            if n.op == "null":
                shapes[id(n)] = [var_shape(n)]
                continue
            ins = [shapes[id(s)][i] for s, i in n.inputs]
            outs, ins_new = _infer(n, ins)
            for (s, i), shp in zip(n.inputs, ins_new):       # parameter shapes deduced from the data shape
                if s.op == "null" and shapes[id(s)][0] is None and shp is not None:
                    shapes[id(s)] = [tuple(shp)]
            shapes[id(n)] = outs
        order = self._topo()
        args = [shapes[id(n)][0] for n in order if n.op == "null" and not n.is_aux]
        auxs = [shapes[id(n)][0] for n in order if n.op == "null" and n.is_aux]
        outs = [shapes[id(n)][i] for n, i in self._heads]
        return args, outs, auxs


def Group(symbols):
    heads = []
    for s in symbols:
        heads.extend(s._heads)
    return Symbol(heads)


def var(name, attr=None, shape=None, lr_mult=None, wd_mult=None, dtype=None, init=None, stype=None, **kwargs):
    attrs = dict(attr or {})
    if shape is not None:
        attrs["__shape__"] = _attr_str(tuple(shape))
    if lr_mult is not None:
        attrs["__lr_mult__"] = str(lr_mult)
    if wd_mult is not None:
        attrs["__wd_mult__"] = str(wd_mult)
    if dtype is not None:
        attrs["__dtype__"] = "0" if str(dtype) in ("float32", "<class 'numpy.float32'>", "0") else str(dtype)
    if init is not None:
        attrs["__init__"] = _attr_str(init)
    for k, v in kwargs.items():
        if k.startswith("__") and k.endswith("__"):
            attrs[k] = _attr_str(v)
        else:
            raise ValueError("Attribute name=%s is not supported" % k)
    return Symbol([(_Node("null", name, attrs, []), 0)])


Variable = var

# operator -> (ordered input names, aux input names, name hint)
_OPS = {
    "Convolution": (["data", "weight", "bias"], [], "convolution"),
    "Deconvolution": (["data", "weight", "bias"], [], "deconvolution"),
    "FullyConnected": (["data", "weight", "bias"], [], "fullyconnected"),
    "BatchNorm": (["data", "gamma", "beta", "moving_mean", "moving_var"], ["moving_mean", "moving_var"], "batchnorm"),
    "BatchNorm_v1": (["data", "gamma", "beta", "moving_mean", "moving_var"], ["moving_mean", "moving_var"], "batchnorm_v1"),
    "Activation": (["data"], [], "activation"),
    "identity": (["data"], [], "identity"),
    "Cast": (["data"], [], "cast"),
    "LeakyReLU": (["data"], [], "leakyrelu"),
    "Pooling": (["data"], [], "pooling"),
    "Flatten": (["data"], [], "flatten"),
    "Dropout": (["data"], [], "dropout"),
    "SoftmaxOutput": (["data", "label"], [], "softmaxoutput"),
    "elemwise_add": (["lhs", "rhs"], [], "_plus"),
    "elemwise_mul": (["lhs", "rhs"], [], "_mul"),
    "broadcast_add": (["lhs", "rhs"], [], "broadcast_add"),
    "broadcast_mul": (["lhs", "rhs"], [], "broadcast_mul"),
    "_plus_scalar": (["data"], [], "_plusscalar"),
    "_mul_scalar": (["data"], [], "_mulscalar"),
    "add_n": (None, [], "add_n"),           # variadic
    "Concat": (None, [], "concat"),         # variadic
    "_contrib_BroadcastScale": (["data", "scaler"], [], "broadcastscale"),
    # fork-only C++ operators (source absent): input slots assumed to mirror the Python twins (minmax / alpha
    # auxiliary, PACT's gamma an argument)
    "_contrib_Quantization_int8": (["data", "minmax"], ["minmax"], "quantization_int8"),
    "_contrib_GDRQ": (["data", "alpha"], ["alpha"], "gdrq"),
    "_contrib_PACT": (["data", "gamma"], [], "pact"),
    "_contrib_DoReFa": (["data"], [], "dorefa"),
}
_ALIASES = {"ElementWiseSum": "add_n", "concat": "Concat", "Concat": "Concat", "flatten": "Flatten",
            "relu": "Activation"}


def _prop_for(op_type, attrs):
    for reg in _REGISTRIES:
        if op_type in reg:
            return reg[op_type](**{k: _attr_str(v) for k, v in attrs.items()})
    raise KeyError("Custom op_type %r is not registered with the shim (set_registries)" % op_type)


def _output_names(node):
    if node.op in ("BatchNorm", "BatchNorm_v1") and node.num_outputs == 3:
        return ["output", "mean", "var"]
    return ["output"] * node.num_outputs


def _optional_inputs_present(op, attrs):
    """which declared inputs exist given the attributes (bias only without no_bias)."""
    names, _, _ = _OPS[op]
    if op in ("Convolution", "Deconvolution", "FullyConnected"):
        default_no_bias = op == "Deconvolution"     # [upstream] Deconvolution defaults to no_bias=True
        if _bool(attrs.get("no_bias"), default_no_bias):
            return [n for n in names if n != "bias"]
    return list(names)


def _make(op, args, kwargs, name, hint=None):
    names, aux_names, default_hint = _OPS[op]
    attrs, sym_kwargs = {}, {}
    for k, v in kwargs.items():
        if isinstance(v, Symbol):
            sym_kwargs[k] = v
        elif v is not None:
            attrs[k] = _attr_str(v)
    name = name or _auto_name(hint or default_hint)
    pos = [a for a in args if a is not None]
    assert all(isinstance(a, Symbol) for a in pos), "positional operator inputs must be symbols"
    inputs = []
    if names is None:                                    # variadic: add_n / Concat
        for a in pos:
            inputs.append(a._heads[0])
        attrs.setdefault("num_args", str(len(pos)))
    else:
        wanted = _optional_inputs_present(op, attrs)
        given = dict(zip(wanted, pos))
        for k, v in sym_kwargs.items():
            if k not in names:
                raise TypeError("%s got an unexpected symbol input %r" % (op, k))
            given[k] = v
        for n in wanted:
            if n in given:
                inputs.append(given[n]._heads[0])
            else:                                        # auto-created parameter / aux variable
                inputs.append((_Node("null", "%s_%s" % (name, n), {}, [], is_aux=(n in aux_names)), 0))
        for n in aux_names:                              # a variable passed for an aux slot is an auxiliary state
            if n in given and given[n]._heads[0][0].op == "null":
                given[n]._heads[0][0].is_aux = True
    nout = 3 if (op in ("BatchNorm", "BatchNorm_v1") and _bool(attrs.get("output_mean_var"))) else 1
    return Symbol([(_Node(op, name, attrs, inputs, nout), 0)])


def Custom(*args, **kwargs):
    name = kwargs.pop("name", None)
    op_type = kwargs.pop("op_type")
    attrs, sym_kwargs = {}, {}
    for k, v in kwargs.items():
        if isinstance(v, Symbol):
            sym_kwargs[k] = v
        elif v is not None:
            attrs[k] = _attr_str(v)
    prop = _prop_for(op_type, attrs)
    name = name or _auto_name("custom")
    arg_names, aux_names = list(prop.list_arguments()), list(prop.list_auxiliary_states())
    # positional inputs run over arguments followed by auxiliary states [upstream custom.cc: FListInputNames]
    given = dict(zip(arg_names + aux_names, [a for a in args if a is not None]))
    given.update(sym_kwargs)
    inputs = []
    for n in arg_names:
        inputs.append(given[n]._heads[0] if n in given else (_Node("null", "%s_%s" % (name, n), {}, []), 0))
    for n in aux_names:
        if n in given:
            node = given[n]._heads[0][0]
            node.is_aux = True
            inputs.append((node, 0))
        else:
            inputs.append((_Node("null", "%s_%s" % (name, n), {}, [], is_aux=True), 0))
    attrs["op_type"] = op_type
    node = _Node("Custom", name, attrs, inputs, len(prop.list_outputs()))
    node.prop = prop
    return Symbol([(node, 0)])


def _op_factory(op):
    def fn(*args, **kwargs):
        name = kwargs.pop("name", None)
        kwargs.pop("attr", None)
        return _make(op, args, kwargs, name)
    fn.__name__ = op
    return fn


class _Contrib(object):
    def __getattr__(self, item):
        op = "_contrib_" + item
        if op not in _OPS:
            raise AttributeError("mx.sym.contrib.%s is not modelled by the shim" % item)
        return _op_factory(op)


contrib = _Contrib()


class _Internal(object):
    def __getattr__(self, item):
        if item not in _OPS:
            raise AttributeError("mx.sym._internal.%s is not modelled by the shim" % item)
        return _op_factory(item)


_internal = _Internal()


def __getattr__(item):       # module-level: mx.sym.<Operator>
    op = _ALIASES.get(item, item)
    if op in _OPS and not op.startswith("_contrib_"):
        return _op_factory(op)
    raise AttributeError("mx.sym.%s is not modelled by the shim" % item)


# ---- shape inference ------------------------------------------------------------------------------------------
def _conv_out(n, k, s, p, d):
    return (n + 2 * p - d * (k - 1) - 1) // s + 1


def _infer(node, ins):
    """-> ([output shapes], [input shapes, with deduced parameter shapes filled in])"""
    op, a = node.op, node.attrs
    x = ins[0] if ins else None
    if op == "Custom":
        prop = node.prop
        nargs = len(prop.list_arguments())
        in_s, out_s, aux_s = prop.infer_shape([list(s) if s is not None else None for s in ins[:nargs]])
        return [tuple(s) for s in out_s], [tuple(s) for s in in_s] + [tuple(s) for s in aux_s]
    if op in ("Convolution", "Deconvolution"):
        k = _tuple(a.get("kernel"))
        s = _tuple(a.get("stride"), (1,) * len(k)) or (1,) * len(k)
        p = _tuple(a.get("pad"), (0,) * len(k)) or (0,) * len(k)
        d = _tuple(a.get("dilate"), (1,) * len(k)) or (1,) * len(k)
        f, g = int(a["num_filter"]), int(a.get("num_group", 1))
        if op == "Convolution":
            out = (x[0], f) + tuple(_conv_out(x[2 + i], k[i], s[i], p[i], d[i]) for i in range(len(k)))
            w = (f, x[1] // g) + k
        else:
            out = (x[0], f) + tuple((x[2 + i] - 1) * s[i] - 2 * p[i] + d[i] * (k[i] - 1) + 1 for i in range(len(k)))
            w = (x[1], f // g) + k
        new = [x, w] + ([(f,)] if len(ins) > 2 else [])
        return [out], new
    if op == "FullyConnected":
        h = int(a["num_hidden"])
        flat = 1
        for v in x[1:]:
            flat *= v
        if not _bool(a.get("flatten"), True):
            return [tuple(x[:-1]) + (h,)], [x, (h, x[-1])] + ([(h,)] if len(ins) > 2 else [])
        return [(x[0], h)], [x, (h, flat)] + ([(h,)] if len(ins) > 2 else [])
    if op in ("BatchNorm", "BatchNorm_v1"):
        c = (x[int(a.get("axis", 1))],)
        outs = [x] + ([c, c] if node.num_outputs == 3 else [])
        return outs, [x, c, c, c, c]
    if op == "Pooling":
        if _bool(a.get("global_pool")):
            return [tuple(x[:2]) + (1,) * (len(x) - 2)], ins
        k = _tuple(a.get("kernel"))
        s = _tuple(a.get("stride"), (1,) * len(k)) or (1,) * len(k)
        p = _tuple(a.get("pad"), (0,) * len(k)) or (0,) * len(k)
        full = a.get("pooling_convention", "valid") == "full"
        dims = []
        for i in range(len(k)):
            num = x[2 + i] + 2 * p[i] - k[i]
            dims.append((-(-num // s[i]) if full else num // s[i]) + 1)
        return [tuple(x[:2]) + tuple(dims)], ins
    if op == "Flatten":
        flat = 1
        for v in x[1:]:
            flat *= v
        return [(x[0], flat)], ins
    if op == "Concat":
        dim = int(a.get("dim", 1))
        out = list(ins[0])
        out[dim] = sum(s[dim] for s in ins)
        return [tuple(out)], ins
    if op == "SoftmaxOutput":
        return [x], [x, (x[0],)]
    if op == "_contrib_BroadcastScale":
        return [x], ins
    if op in ("_contrib_Quantization_int8", "_contrib_GDRQ", "_contrib_PACT"):
        return [x], [x, (1,)]
    # element-wise / same-shape operators
    return [x], [s if s is not None else x for s in ins]
