"""Drive one operator instance (reference-over-shim, oracle, or the CUDA op) through a golden case.

``make_inputs`` builds every externally supplied array of a case from a seeded RNG (pure numpy).
``drive`` feeds them to an operator that speaks the MXNet CustomOp protocol and returns every array the
operator wrote, keyed ``s<step>_...``.  The same function produces the fixtures and replays them.
"""
import numpy as np

from .cases import build_step_inputs, case_seed

F = np.float32
SENTINEL = F(777.0)


def out_shape_of(case, prop):
    in_shapes = [list(s[1]) for s in case["inputs"]]
    _, out_shapes, aux_shapes = prop.infer_shape(in_shapes)
    return [tuple(s) for s in out_shapes], [tuple(s) for s in aux_shapes]


def make_inputs(case, out_shapes):
    rng = np.random.default_rng(case_seed(case))
    d = {}
    for k, (is_train, req, do_bwd) in enumerate(case["steps"]):
        for i, a in enumerate(build_step_inputs(case, rng, k)):
            d["s%d_in%d" % (k, i)] = a
        if req == "add":
            d["s%d_out_init" % k] = rng.standard_normal(out_shapes[0]).astype(F)
        else:
            d["s%d_out_init" % k] = np.full(out_shapes[0], SENTINEL, dtype=F)
        if do_bwd:
            d["s%d_og" % k] = rng.standard_normal(out_shapes[0]).astype(F)
            for i, spec in enumerate(case["inputs"]):
                if req == "add":
                    d["s%d_ig_init%d" % (k, i)] = rng.standard_normal(spec[1]).astype(F)
                else:
                    d["s%d_ig_init%d" % (k, i)] = np.full(spec[1], SENTINEL, dtype=F)
    return d


def drive(case, inputs, op, aux_shapes, to_arr, to_np):
    """Returns dict of results.  ``to_arr(np_array) -> framework array`` (fresh, writable);
    ``to_np(framework array) -> np.ndarray``."""
    res = {}
    n_in = len(case["inputs"])
    aux = [to_arr(np.full(shp, case["aux_init"][j], dtype=F)) for j, shp in enumerate(aux_shapes)]
    for k, (is_train, req, do_bwd) in enumerate(case["steps"]):
        in_data = [to_arr(inputs["s%d_in%d" % (k, i)].copy()) for i in range(n_in)]
        out_data = [to_arr(inputs["s%d_out_init" % k].copy())]
        try:
            op.forward(is_train, [req], in_data, out_data, aux)
        except NameError as e:  # GDRQ_Fold_BN at inference (fold_bn_v1_gdrq.py:67)
            res["s%d_raises" % k] = np.array([1], dtype=np.int32)
            continue
        res["s%d_out" % k] = to_np(out_data[0])
        for i in range(n_in):
            res["s%d_in%d_after" % (k, i)] = to_np(in_data[i])
        for j in range(len(aux)):
            res["s%d_aux%d" % (k, j)] = to_np(aux[j])
        if do_bwd:
            out_grad = [to_arr(inputs["s%d_og" % k].copy())]
            in_grad = [to_arr(inputs["s%d_ig_init%d" % (k, i)].copy()) for i in range(n_in)]
            op.backward([req] * n_in, out_grad, in_data, out_data, in_grad, aux)
            for i in range(n_in):
                res["s%d_ig%d" % (k, i)] = to_np(in_grad[i])
            for j in range(len(aux)):
                res["s%d_aux%d_after_bwd" % (k, j)] = to_np(aux[j])
    return res
