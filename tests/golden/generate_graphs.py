"""Generate tests/golden/graphs.json by running the REFERENCE's own symbol builders (symbol/quant_ops.py,
symbol/int8_api.py, symbol/fold_bn_v1_gdrq.py, core/graph_optimize.py -- imported by path, unmodified) over the shim's
``mx.sym`` graph recorder on the cases of ``graph_cases.py``.  Build container only (needs /root/reference):

    python -m tests.golden.generate_graphs
"""
import importlib
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("B2Q_REFERENCE", "/root/reference")


def reference_namespace():
    """The reference's builders, importable without executing its package __init__ files (which pull in every model)."""
    sys.path.insert(0, ROOT)
    import oracle.mxshim as shim
    mx = shim.install()
    for pkg, sub in (("_b2q_refpkg_symbol", "symbol"), ("_b2q_refpkg_core", "core")):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(REF, sub)]
            sys.modules[pkg] = m
    quant_ops = importlib.import_module("_b2q_refpkg_symbol.quant_ops")
    int8_api = importlib.import_module("_b2q_refpkg_symbol.int8_api")
    fold = importlib.import_module("_b2q_refpkg_symbol.fold_bn_v1_gdrq")
    go = importlib.import_module("_b2q_refpkg_core.graph_optimize")
    import torch
    ns = types.SimpleNamespace(mx=mx, quant_conv=quant_ops.quant_conv, quant_fc=quant_ops.quant_fc,
                               GDRQ_fold_bn=fold.GDRQ_fold_bn, create_quant_node=go.create_quant_node,
                               attach_quantize_node=go.attach_quantize_node, merge_bn=go.merge_bn, fix_bn=go.fix_bn,
                               to_array=lambda a: shim.NDArray(torch.from_numpy(np.array(a, dtype=np.float32))),
                               to_numpy=lambda a: a.asnumpy())
    for name in ("clipgrad_quant_conv", "clipgrad_quant_fc", "clipgrad_quant_deconv", "clipgrad_quant_data",
                 "clipgrad_quant_add", "clipgrad_quant_concat", "quant_conv_cxx", "quant_fc_cxx", "quant_deconv_cxx",
                 "quant_add_cxx", "quant_concat_cxx"):
        setattr(ns, name, getattr(int8_api, name))
    mx.sym.set_registries(shim.REGISTRY)
    return ns, shim


def resnet50_inventory(ns_modules=None):
    """Quantization nodes of the reference's own resnet_int8 symbol (symbol/resnet_int8.py:69-131, depth 50, batch 256):
    (node name, is_weight, input shape) in graph order -- pins b200quant.workloads.resnet50_nodes."""
    mod = importlib.import_module("_b2q_refpkg_symbol.resnet_int8")
    import io
    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        sym = mod.resnet_int8(units=[3, 4, 6, 3], num_stage=4, filter_list=[64, 256, 512, 1024, 2048], num_classes=1000,
                              data_type="float32", bottle_neck=True, dataset_type="imagenet",
                              dict_shapes={"data": (256, 3, 224, 224)})
    g = json.loads(sym.tojson())
    internals = sym.get_internals()
    names = internals.list_outputs()
    _, shapes, _ = internals.infer_shape(data=(256, 3, 224, 224))
    shape_of = dict(zip(names, shapes))
    inv = []
    for n in g["nodes"]:
        if n["op"] == "_contrib_Quantization_int8":
            src = g["nodes"][n["inputs"][0][0]]
            key = src["name"] if src["op"] == "null" else src["name"] + "_output"
            inv.append([n["name"], n["attrs"]["is_weight"] == "True", list(shape_of[key])])
    return inv


def main():
    from tests.golden import graph_cases as gc
    ns, shim = reference_namespace()
    out = {}
    for case in gc.CASES:
        ns.mx.sym.reset_names()
        try:
            out[case.__name__] = gc.normalize(case(ns))
        except Exception as e:   # the reference's own failure modes are part of the contract (WNQ: UnboundLocalError)
            out[case.__name__] = {"raises": type(e).__name__}
        print("%-36s %s" % (case.__name__, out[case.__name__].get("raises") or "%d nodes" % len(out[case.__name__]["nodes"])))
    ns.mx.sym.reset_names()
    sym, args, auxs = gc.merge_bn_arrays_case(ns, np)
    out["merge_bn_arrays"] = {"graph": gc.normalize(sym),
                              "args": {k: [list(v.shape), np.asarray(v, np.float64).ravel().tolist()] for k, v in args.items()},
                              "auxs": {k: [list(v.shape), np.asarray(v, np.float64).ravel().tolist()] for k, v in auxs.items()}}
    out["resnet50_int8_inventory"] = resnet50_inventory()
    print("resnet50 inventory: %d quantization nodes" % len(out["resnet50_int8_inventory"]))
    with open(os.path.join(HERE, "graphs.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
