"""Symbol-graph cases shared by the golden generator (run on the REFERENCE's builders) and tests/test_symbol_builders.py
(run on this repository's builders).  Every case is ``fn(ns) -> symbol`` where ``ns`` offers the same names on both sides:
``mx`` (the mxnet stand-in), the builders of symbol/quant_ops.py, symbol/int8_api.py, symbol/fold_bn_v1_gdrq.py and the
four entry points of core/graph_optimize.py."""


class Setting(dict):
    """config.quantize_setting entry with attribute access, like the reference's EasyDict (config/edict_config.py:158)."""
    __getattr__ = dict.get


def V2(is_weight):
    return Setting(quantize_op_name="Quantization_int8", init_value=0,
                   attrs={"nbits": "8", "quant_mode": "minmax", "is_weight": str(is_weight), "is_weight_perchannel": "False",
                          "delay_quant": "0", "ema_decay": "0.99", "grad_mode": "ste", "fix_act_scale": "False"})


def GDRQ(name, is_weight, init=None):
    return Setting(quantize_op_name=name, init_value=init,
                   attrs={"nbits": "4", "fix_alpha": "False", "group_size": "-1", "is_weight": str(is_weight),
                          "lamda": "0.001", "delay_quant": "0", "ktimes": "3"})


SETTINGS = {
    "Quantization_int8": V2(False),
    "QIL": Setting(quantize_op_name="QIL", init_value=None, attrs={"is_weight": "False", "fix_gamma": "True", "nbits": "4"}),
    "DoReFa_PY": Setting(quantize_op_name="DoReFa_PY", attrs={"nbits": "4"}),
    "DoReFa_CXX": Setting(quantize_op_name="DoReFa_CXX", attrs={"nbits": "4"}),
    "PACT": Setting(quantize_op_name="PACT", init_value=6.0, attrs={"nbits": "4"}),
    "PACT_CXX": Setting(quantize_op_name="PACT_CXX", init_value=None, attrs={"nbits": "4"}),
    "GDRQ": GDRQ("GDRQ", False, 0.5),
    "GDRQ_CXX": GDRQ("GDRQ_CXX", True),
    "WNQ": Setting(quantize_op_name="WNQ", attrs={"nbits": "4", "is_perchannel": "False"}),
}
SHAPES = {"data": (2, 3, 16, 16)}


def _small_net(mx):
    """data -> conv0 -> bn0 -> relu0 -> pool0 -> {conv1, conv2 (same input)} -> concat -> conv3 -> (+ pool0 via 1x1 sc) ->
    deconv -> flatten -> fc."""
    data = mx.sym.Variable("data")
    c0 = mx.sym.Convolution(data=data, num_filter=8, kernel=(3, 3), stride=(1, 1), pad=(1, 1), no_bias=True, name="conv0")
    b0 = mx.sym.BatchNorm(data=c0, eps=1e-5, momentum=0.9, fix_gamma=False, use_global_stats=True, name="bn0")
    r0 = mx.sym.Activation(data=b0, act_type="relu", name="relu0")
    p0 = mx.sym.Pooling(data=r0, kernel=(2, 2), stride=(2, 2), pool_type="max", name="pool0")
    c1 = mx.sym.Convolution(data=p0, num_filter=4, kernel=(1, 1), no_bias=True, name="conv1")
    c2 = mx.sym.Convolution(data=p0, num_filter=4, kernel=(3, 3), pad=(1, 1), no_bias=False, name="conv2")
    cat = mx.sym.Concat(c1, c2, dim=1, name="cat0")
    b1 = mx.sym.BatchNorm(data=cat, eps=2e-5, momentum=0.9, fix_gamma=False, use_global_stats=False, name="bn1")
    c3 = mx.sym.Convolution(data=b1, num_filter=8, kernel=(3, 3), pad=(1, 1), no_bias=True, name="conv3")
    s = mx.sym.elemwise_add(c3, p0, name="add0")
    s2 = mx.sym.add_n(s, c3, name="addn0")
    d = mx.sym.Deconvolution(data=s2, num_filter=4, kernel=(2, 2), stride=(2, 2), pad=(0, 0), no_bias=True, name="deconv0")
    f = mx.sym.Flatten(data=d, name="flat0")
    fc = mx.sym.FullyConnected(data=f, num_hidden=10, name="fc1")
    return mx.sym.SoftmaxOutput(data=fc, name="softmax")


def _shape_dict(sym, batch_shapes):
    args, _, auxs = sym.infer_shape(**batch_shapes)
    d = dict(zip(sym.list_arguments(), args))
    d.update(zip(sym.list_auxiliary_states(), auxs))
    return d


def case_quant_ops(ns):
    mx = ns.mx
    data = mx.sym.Variable("data")
    c = ns.quant_conv("conv0", data, 8, (3, 3), (1, 1), pad=(1, 1))
    r = mx.sym.Activation(data=c, act_type="relu", name="relu0")
    c1 = ns.quant_conv("conv1", r, 16, (3, 3), (2, 2), pad=(1, 1), num_group=2, delay_quant=3, is_weight_perchannel=True)
    p = mx.sym.Pooling(data=c1, global_pool=True, kernel=(7, 7), pool_type="avg", name="pool1")
    f = mx.sym.Flatten(data=p, name="flat")
    return ns.quant_fc("fc1", f, 10, delay_quant=1)


def case_int8_clipgrad(ns):
    mx = ns.mx
    data = mx.sym.Variable("data")
    c = ns.clipgrad_quant_conv("conv0", data, 8, (3, 3), (1, 1), pad=(1, 1), no_bias=True, dict_shapes=SHAPES,
                               is_weight_perchannel=True, lr_mult=2.0)
    r = mx.sym.Activation(data=c, act_type="relu", name="relu0")
    a = ns.clipgrad_quant_add("res0", r, c, ema_decay=0.9)
    c1 = ns.clipgrad_quant_conv("conv1", a, 4, (1, 1), (1, 1), no_bias=False, dict_shapes=SHAPES)
    cat = ns.clipgrad_quant_concat("cat0", [c1, a], dim=1)
    d = ns.clipgrad_quant_deconv("up0", cat, (2, 2), (2, 2), (0, 0), 6, dict_shapes=SHAPES)
    f = mx.sym.Flatten(data=d, name="flat")
    return ns.clipgrad_quant_fc("fc1", f, 10, dict_shapes=SHAPES, delay_quant=1)


def case_int8_clipgrad_data(ns):
    """the reference's clipgrad_quant_data omits is_weight_perchannel, so its Prop eval()s the bool default and raises
    TypeError (clip_grad_quantization_int8.py:77); this package's Prop accepts the default."""
    mx = ns.mx
    data = mx.sym.Variable("data")
    return ns.clipgrad_quant_data("in0", data, delay_quant=2)


def case_int8_cxx(ns):
    mx = ns.mx
    data = mx.sym.Variable("data")
    c = ns.quant_conv_cxx("conv0", data, 8, (3, 3), (1, 1), pad=(1, 1), no_bias=True, dict_shapes=SHAPES, grad_mode="clip")
    a = ns.quant_add_cxx("res0", c, c)
    cat = ns.quant_concat_cxx("cat0", [a, c], dim=1)
    d = ns.quant_deconv_cxx("up0", cat, (2, 2), (2, 2), (0, 0), 6, dict_shapes=SHAPES)
    f = mx.sym.Flatten(data=d, name="flat")
    return ns.quant_fc_cxx("fc1", f, 10, dict_shapes=SHAPES, workspace=256)


def case_foldbn(ns):
    mx = ns.mx
    data = mx.sym.Variable("data")
    y = ns.GDRQ_fold_bn("stage1", data, quant_mod="minmax", is_weight_perchannel=True, delay_quant=5, ema_decay=0.95,
                        num_filter=8, kernel=(3, 3), stride=(2, 2), pad=(1, 1), num_group=1, eps=1e-3, momentum=0.8,
                        dict_shapes=SHAPES, w_lr_mult=0.5)
    return mx.sym.Activation(data=y, act_type="relu", name="relu")


def make_create_case(name):
    def case(ns):
        var = ns.mx.sym.Variable("relu0")
        return ns.create_quant_node(var, SETTINGS[name])
    case.__name__ = "case_create_" + name
    return case


def case_attach_default(ns):
    net = _small_net(ns.mx)
    return ns.attach_quantize_node(net, _shape_dict(net, SHAPES), V2(True), V2(False))


def case_attach_all_ops(ns):
    net = _small_net(ns.mx)
    return ns.attach_quantize_node(net, _shape_dict(net, SHAPES), GDRQ("GDRQ", True), SETTINGS["PACT"],
                                   quantized_op=("Convolution", "FullyConnected", "Deconvolution", "Concat", "Pooling",
                                                 "add_n", "elemwise_add"))


def case_attach_skip(ns):
    net = _small_net(ns.mx)
    return ns.attach_quantize_node(net, _shape_dict(net, SHAPES), V2(True), SETTINGS["GDRQ"],
                                   quantized_op=("Convolution", "FullyConnected", "Pooling"),
                                   skip_quantize_counts={"Convolution": 2, "Pooling": 1})


def case_fix_bn(ns):
    return ns.fix_bn(_small_net(ns.mx))


def case_merge_bn_symbol_only(ns):
    out, _, _ = ns.merge_bn(_small_net(ns.mx), None, None, True)
    return out


CASES = [case_quant_ops, case_int8_clipgrad, case_int8_clipgrad_data, case_int8_cxx, case_foldbn, case_attach_default, case_attach_all_ops,
         case_attach_skip, case_fix_bn, case_merge_bn_symbol_only] + [make_create_case(n) for n in SETTINGS]


def normalize(sym):
    """JSON graph -> comparable structure (node order kept: it is the construction order both sides must share)."""
    import json
    g = json.loads(sym.tojson())
    nodes = g["nodes"]
    out = []
    for n in nodes:
        out.append({"op": n["op"], "name": n["name"], "attrs": dict(sorted(n.get("attrs", {}).items())),
                    "inputs": [[nodes[e[0]]["name"], e[1]] for e in n["inputs"]]})
    return {"nodes": out, "heads": [[nodes[e[0]]["name"], e[1]] for e in g["heads"]],
            "arguments": list(sym.list_arguments()), "aux": list(sym.list_auxiliary_states())}


def merge_bn_arrays_case(ns, np):
    """merge_bn with parameter arrays: returns the folded arrays (as numpy) next to the graph."""
    mx = ns.mx
    net = _small_net(mx)
    shapes = _shape_dict(net, SHAPES)
    rng = np.random.default_rng(5)
    make = ns.to_array
    args = {k: make(rng.uniform(0.5, 1.5, shapes[k]).astype(np.float32)) for k in net.list_arguments() if k not in ("data", "softmax_label")}
    auxs = {k: make(rng.uniform(0.5, 1.5, shapes[k]).astype(np.float32)) for k in net.list_auxiliary_states()}
    out, args, auxs = ns.merge_bn(net, args, auxs, False)
    back = ns.to_numpy
    return out, {k: back(v) for k, v in args.items()}, {k: back(v) for k, v in auxs.items()}
