"""Known-answer vectors from the reference's own NumPy simulators (``simulate_GDRQ``
core/operator/GDRQ.py:8-48, ``simulate_PACT`` / ``simulate_DoReFa`` core/operator/PACT.py:10-23, ``simulate_wnq`` /
``simulate_wnq_backword`` core/operator/WNQ.py:5-38).  They are the only known-answer helpers the reference holds.

Run in the build container only (needs /root/reference):

    python -m tests.golden.generate_simulators        # from the repo root

The simulators are float64 NumPy with ``np.round`` (half to even); the operators are float32 with half-away-from-zero
rounding, so every vector also stores ``near_tie``: the elements whose pre-rounding value lies within 1e-3 of a
half-integer, which the tests leave out.  ``simulate_GDRQ`` hard-wires ``ktimes = 2``, and for activations
``alpha = 0.5, lamda = 0.01`` (:13,16); for weights it quantizes to ``nbits - 1`` bits (:9-10) -- the tests configure the
operator accordingly.
"""
import os

import numpy as np

from tests.golden.generate import HERE, REF, load_reference


def _near_tie(t):
    frac = np.abs(t - np.floor(t) - 0.5)
    return frac < 1e-3


def main():
    load_reference()   # installs the mxnet shim (the modules start with `import mxnet`)
    import importlib.util

    def ref_module(rel):
        spec = importlib.util.spec_from_file_location("_b2q_sim_" + os.path.basename(rel)[:-3], os.path.join(REF, rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    gd = ref_module("core/operator/GDRQ.py")
    pa = ref_module("core/operator/PACT.py")
    wn = ref_module("core/operator/WNQ.py")
    rng = np.random.default_rng(5)
    out = {}

    # PACT: simulate_PACT clips to [0, gamma]; the operator only clips from above (PACT.py:118), so use post-ReLU data
    x = np.maximum(rng.standard_normal((4, 8, 6, 6)) * 3.0, 0).astype(np.float32)
    gamma, nbits = 4.0, 4
    out["pact_x"], out["pact_gamma"], out["pact_nbits"] = x, np.float32(gamma), np.int32(nbits)
    out["pact_y"] = pa.simulate_PACT(x.astype(np.float64), gamma, nbits)
    out["pact_near_tie"] = _near_tie(np.clip(x.astype(np.float64), 0, gamma) / (gamma / (2 ** nbits - 1)))

    # DoReFa weights
    w = (rng.standard_normal((8, 4, 3, 3)) * 0.7).astype(np.float32)
    nbits = 4
    out["dorefa_x"], out["dorefa_nbits"] = w, np.int32(nbits)
    out["dorefa_y"] = pa.simulate_DoReFa(w.astype(np.float64), nbits)
    t = np.tanh(w.astype(np.float64))
    out["dorefa_near_tie"] = _near_tie((t / (2 * np.max(np.abs(t))) + 0.5) * (2 ** nbits - 1))

    # WNQ forward / backward, per tensor and per channel
    w = (rng.standard_normal((6, 5, 3, 3)) * 0.2).astype(np.float32)
    dy = rng.standard_normal(w.shape).astype(np.float32)
    L = 2 ** 4 - 1
    out["wnq_x"], out["wnq_dy"], out["wnq_level"] = w, dy, np.int32(L)
    for tag, pc in (("pt", False), ("pc", True)):
        out["wnq_y_" + tag] = wn.simulate_wnq(w.astype(np.float64), pc, L)
        out["wnq_dx_" + tag] = wn.simulate_wnq_backword(w.astype(np.float64), dy.astype(np.float64), pc, L)
        m = np.max(np.abs(w.astype(np.float64)), axis=(1, 2, 3), keepdims=True) if pc else np.max(np.abs(w))
        out["wnq_near_tie_" + tag] = _near_tie(w.astype(np.float64) / m * L)

    # GDRQ: weights (simulator nbits 5 == operator nbits 4) and activations (alpha 0.5, lamda 0.01), whole and grouped
    for tag, shape, is_weight, group in (("w", (8, 4, 3, 3), True, -1), ("wg", (8, 4, 3, 3), True, 2),
                                         ("a", (3, 8, 5, 5), False, -1), ("ag", (3, 8, 5, 5), False, 4)):
        x = (rng.standard_normal(shape) * 0.3).astype(np.float32)
        sim_bits = 5 if is_weight else 4
        y = gd.simulate_GDRQ(x.astype(np.float64), sim_bits, group, is_weight)
        out["gdrq_%s_x" % tag] = x
        out["gdrq_%s_y" % tag] = y
        out["gdrq_%s_meta" % tag] = np.array([4, group, int(is_weight)], np.int32)   # operator nbits, group, is_weight
        # pre-rounding codes, recomputed from the simulator's own output grid: y / unit is integral, so locate near ties
        # from the clipped input instead
        x64 = x.astype(np.float64)
        if group == -1:
            thr = 2 * np.mean(np.abs(x64))
            if not is_weight:
                thr = 0.5 + 0.01 * (0.5 - thr)
            c = np.clip(x64, -thr, thr)
            t = c / (thr / (2 ** 4 - 1))
        else:
            v = x64 if is_weight else np.swapaxes(x64, 0, 1)
            r = v.reshape((v.shape[0] // group, -1))
            thr = 2 * np.mean(np.abs(r), axis=1, keepdims=True)
            if not is_weight:
                thr = 0.5 + 0.01 * (0.5 - thr)
            c = np.where(np.abs(r) <= thr, r, thr * np.sign(r))
            t = (c / (thr / (2 ** 4 - 1))).reshape(v.shape)
            if not is_weight:
                t = np.swapaxes(t, 0, 1)
        out["gdrq_%s_near_tie" % tag] = _near_tie(t)

    np.savez_compressed(os.path.join(HERE, "simulators.npz"), **out)
    print("wrote simulators.npz with %d arrays" % len(out))


if __name__ == "__main__":
    main()
