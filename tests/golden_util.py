"""Load golden fixtures and compare a driven operator against them."""
import json
import os

import numpy as np

from tests.golden.cases import CASES, CASE_BY_ID  # noqa: F401
from tests.golden.driver import drive

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

with open(os.path.join(GOLDEN, "manifest.json")) as _f:
    MANIFEST = json.load(_f)

# arrays whose value goes through library code / autograd replay on the reference side: tolerance, not bits
_LOOSE_OPS = {"DoReFa_PY", "QIL_PY", "QIL_V2_PY", "QIL_V3_PY", "PACT_PY", "PACT_V2_PY", "WNQ_PY"}


def load(case_id):
    with np.load(os.path.join(GOLDEN, case_id + ".npz")) as z:
        return {k: z[k] for k in z.files}


def bits_equal(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))


def check_against_golden(case, op, to_arr, to_np, exact=True, rtol=1e-6, atol=0.0):
    """Drive ``op`` with the fixture's inputs and compare every recorded array."""
    fx = load(case["id"])
    man = MANIFEST[case["id"]]
    got = drive(case, fx, op, [tuple(s) for s in man["aux_shapes"]], to_arr, to_np)
    assert sorted(got.keys()) == man["keys"], (sorted(got.keys()), man["keys"])
    loose = case["op_type"] in _LOOSE_OPS
    for k in man["keys"]:
        ref, val = fx[k], got[k]
        if k.endswith("_raises"):
            continue
        is_conv_out = case["op_type"] == "GDRQ_Fold_BN" and k.endswith("_out") and \
            not bits_equal(ref, fx[k[:-4] + "_in2"])  # delay branch copies bn_output exactly
        if is_conv_out:
            np.testing.assert_allclose(val, ref, rtol=2e-5, atol=2e-5, err_msg=case["id"] + ":" + k)
        elif loose and ("_ig" in k):
            np.testing.assert_allclose(val, ref, rtol=2e-5, atol=1e-6, err_msg=case["id"] + ":" + k)
        elif loose:
            np.testing.assert_allclose(val, ref, rtol=1e-6, atol=1e-7, err_msg=case["id"] + ":" + k)
        elif exact:
            assert bits_equal(val, ref), "%s:%s differs (max abs %g)" % (
                case["id"], k, float(np.nanmax(np.abs(val.astype(np.float64) - ref.astype(np.float64)))))
        else:
            np.testing.assert_allclose(val, ref, rtol=rtol, atol=atol, err_msg=case["id"] + ":" + k)
    return got
