"""Generate the golden fixtures in this directory by running the REFERENCE's own op classes.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python -m tests.golden.generate            # from the repo root

The reference files are imported by path, unmodified, on top of ``oracle/mxshim`` (an emulation of the
``mxnet`` API surface they use; libmxnet cannot be installed here).  For every case in ``cases.py`` the
reference operator is created through its registered ``CustomOpProp`` (so string-attribute parsing,
``infer_shape`` and ``list_*`` are the reference's), driven through the case's steps, and every array it
writes is stored in ``<case id>.npz``.  ``manifest.json`` records the prop-level contracts.
"""
import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("B2Q_REFERENCE", "/root/reference")

REF_FILES = [
    "symbol/quant_ops.py",
    "symbol/clip_grad_quantization_int8.py",
    "symbol/fold_bn_v1_gdrq.py",
    "core/operator/GDRQ.py",
    "core/operator/PACT.py",
    "core/operator/WNQ.py",
    "core/operator/QIL.py",
    "core/operator/QIL_V2.py",
    "core/operator/QIL_V3.py",
]


def load_reference():
    sys.path.insert(0, ROOT)
    import oracle.mxshim as shim
    shim.install()
    for rel in REF_FILES:
        path = os.path.join(REF, rel)
        name = "_b2q_ref_" + rel.replace("/", "_")[:-3]
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    return shim


def main():
    shim = load_reference()
    import torch
    from tests.golden.cases import CASES
    from tests.golden.driver import drive, make_inputs, out_shape_of

    def to_arr(a):
        return shim.NDArray(torch.from_numpy(np.array(a, dtype=np.float32)))

    def to_np(a):
        return a.asnumpy()

    manifest = {}
    for case in CASES:
        prop = shim.REGISTRY[case["op_type"]](**case["attrs"])
        out_shapes, aux_shapes = out_shape_of(case, prop)
        inputs = make_inputs(case, out_shapes)
        op = prop.create_operator(None, None, None)
        res = drive(case, inputs, op, aux_shapes, to_arr, to_np)
        blob = dict(inputs)
        blob.update(res)
        np.savez_compressed(os.path.join(HERE, case["id"] + ".npz"), **blob)
        manifest[case["id"]] = dict(
            op_type=case["op_type"], attrs=case["attrs"],
            list_arguments=list(prop.list_arguments()), list_outputs=list(prop.list_outputs()),
            list_auxiliary_states=list(prop.list_auxiliary_states()),
            in_shapes=[list(s[1]) for s in case["inputs"]],
            out_shapes=[list(s) for s in out_shapes], aux_shapes=[list(s) for s in aux_shapes],
            keys=sorted(res.keys()))
        print("%-36s %-28s %3d arrays" % (case["id"], case["op_type"], len(res)))
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("wrote %d fixtures to %s" % (len(CASES), HERE))


if __name__ == "__main__":
    main()
