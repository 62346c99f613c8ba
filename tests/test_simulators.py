"""Known-answer check against the reference's own NumPy simulators (the only known-answer helpers it holds, SURVEY.md
section 4): simulate_PACT / simulate_DoReFa (core/operator/PACT.py:10-23), simulate_wnq / simulate_wnq_backword
(core/operator/WNQ.py:5-38), simulate_GDRQ (core/operator/GDRQ.py:8-48).  Vectors: tests/golden/simulators.npz, made by
tests/golden/generate_simulators.py from the reference sources.  The simulators are float64 with np.round (half to
even); the operators are float32, half away from zero: elements within 1e-3 of a rounding tie are left out, everything
else has to agree to float32 accuracy.  Checked for the oracle on CPU and for the CUDA operators on the GPU."""
import os

import numpy as np
import pytest

SIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "simulators.npz")
F = np.float32


def _load():
    with np.load(SIM) as z:
        return {k: z[k] for k in z.files}


def _close(got, want, near_tie, what, rtol=2e-6, atol=2e-6):
    keep = ~near_tie
    assert keep.mean() > 0.9, what
    np.testing.assert_allclose(np.asarray(got, np.float64)[keep], want[keep], rtol=rtol, atol=atol, err_msg=what)


def _run(make, to_arr, to_np):
    """make(op_type, **attrs) -> operator; to_arr / to_np convert to and from the operator's array type."""
    z = _load()
    ones = lambda shape=(1,), v=1.0: to_arr(np.full(shape, v, F))

    # PACT (gamma is a learnable argument)
    op = make("PACT_PY", nbits=str(int(z["pact_nbits"])))
    x = to_arr(z["pact_x"])
    y = to_arr(np.zeros_like(z["pact_x"]))
    op.forward(True, ["write"], [x, to_arr(np.array([z["pact_gamma"]], F))], [y], [])
    _close(to_np(y), z["pact_y"], z["pact_near_tie"], "simulate_PACT")

    # DoReFa
    op = make("DoReFa_PY", nbits=str(int(z["dorefa_nbits"])))
    y = to_arr(np.zeros_like(z["dorefa_x"]))
    op.forward(True, ["write"], [to_arr(z["dorefa_x"])], [y], [])
    _close(to_np(y), z["dorefa_y"], z["dorefa_near_tie"], "simulate_DoReFa", rtol=1e-5, atol=1e-5)

    # WNQ forward + backward (the backward has no rounding: no tie mask needed, but the max element carries a sum)
    for tag, pc in (("pt", False), ("pc", True)):
        op = make("WNQ_PY", nbits="4", is_perchannel=str(pc))
        x, dy = to_arr(z["wnq_x"]), to_arr(z["wnq_dy"])
        y, dx = to_arr(np.zeros_like(z["wnq_x"])), to_arr(np.zeros_like(z["wnq_x"]))
        op.forward(True, ["write"], [x], [y], [])
        _close(to_np(y), z["wnq_y_" + tag], z["wnq_near_tie_" + tag], "simulate_wnq " + tag)
        op.backward(["write"], [dy], [x], [y], [dx], [])
        np.testing.assert_allclose(to_np(dx), z["wnq_dx_" + tag], rtol=2e-5, atol=2e-5,
                                   err_msg="simulate_wnq_backword " + tag)

    # GDRQ: ktimes 2; weights at operator nbits 4 (simulator 5); activations alpha 0.5, lamda 0.01
    for tag in ("w", "wg", "a", "ag"):
        nbits, group, is_weight = (int(v) for v in z["gdrq_%s_meta" % tag])
        x = z["gdrq_%s_x" % tag]
        op = make("GDRQ_PY", nbits=str(nbits), group_size=str(group), is_weight=str(bool(is_weight)), lamda="0.01",
                  ktimes="2")
        channels = x.shape[0] if is_weight else x.shape[1]
        alpha = ones((1 if group == -1 else channels // group,), 0.5)
        y = to_arr(np.zeros_like(x))
        op.forward(True, ["write"], [to_arr(x)], [y], [alpha])
        _close(to_np(y), z["gdrq_%s_y" % tag], z["gdrq_%s_near_tie" % tag], "simulate_GDRQ " + tag)


def test_oracle_matches_reference_simulators():
    from oracle import quant_oracle as qo
    _run(qo.create, lambda a: np.array(a, dtype=F), lambda a: np.array(a))


@pytest.mark.gpu
def test_cuda_operators_match_reference_simulators():
    import torch
    import b200quant

    def make(op_type, **attrs):
        return b200quant.get_prop(op_type)(**attrs).create_operator(None, None, None)

    _run(make, lambda a: torch.from_numpy(np.array(a, dtype=F)).cuda(), lambda t: t.cpu().numpy())
