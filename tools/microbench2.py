"""Micro-benchmark of the kernels beyond the ResNet-50 headline: sum reductions, segmented / grouped sweeps, the fold-BN
weight path, per-channel weights and the second-tier operators, on 2^26-element tensors and the real
weight shapes.  CUDA events, rotating buffers so nothing is served from L2 (6 x 256 MB per role).

    python tools/microbench2.py            -> prints a table, writes gpurun_out/microbench2.json
    bench.py imports run() for the `micro` key of its JSON line (quick=True: fewer repetitions)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def run(torch, ctx, peak_gbs, quick=False, verbose=False):
    import b200quant
    from b200quant import _kernels as K
    from b200quant import _lib

    NB = 4 if quick else 6
    reps_big = 8 if quick else 12
    once = os.environ.get("B2Q_MICRO_ONCE") == "1"     # profiling aid: every line launched twice (ncu captures them)
    rows = {}

    def bench(name, fn, alg_bytes, reps=reps_big):
        if once:
            reps = 1
        for i in range(1 if once else 3):
            fn(i % NB)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            fn(i % NB)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = alg_bytes / ms / 1e6
        rows[name] = {"us": round(ms * 1e3, 2), "gbs": round(gbs, 1), "frac": round(gbs / peak_gbs, 4)}
        if verbose:
            print("%-62s %9.3f ms %8.1f GB/s  %.3f" % (name, ms, gbs, gbs / peak_gbs), flush=True)

    def op(op_type, **attrs):
        return b200quant.get_prop(op_type)(**{k: str(v) for k, v in attrs.items()}).create_operator(None, None, None)

    shape = (256, 64, 64, 64)          # 2^26 elements, 256 MB
    n = 1 << 26
    g = torch.Generator(device="cuda").manual_seed(3)
    xs = [torch.empty(shape, device="cuda").uniform_(-1, 1, generator=g) for _ in range(NB)]
    ys = [torch.empty(shape, device="cuda") for _ in range(NB)]
    dys = [torch.empty(shape, device="cuda").normal_(generator=g) for _ in range(NB)]

    # ---- GDRQ_PY activation, per tensor and grouped (core/operator/GDRQ.py) ----
    for gs in (-1, 16, 1):
        ng = 1 if gs == -1 else 64 // gs
        o = op("GDRQ_PY", nbits=8, group_size=gs, is_weight=False, lamda=0.001, delay_quant=0, fix_alpha=False, ktimes=3)
        alpha = torch.ones(ng, device="cuda")
        bench("GDRQ act fwd group_size=%d (reduce sum + sweep)" % gs,
              lambda i: o.forward(True, ["write"], [xs[i]], [ys[i]], [alpha]), 12 * n)
        bench("GDRQ act bwd group_size=%d (|x|<=alpha mask)" % gs,
              lambda i: o.backward(["write"], [dys[i]], [xs[i]], [ys[i]], [ys[(i + 1) % NB]], [alpha]), 12 * n)

    # ---- grouped pieces in isolation ----
    for gs in (16, 1):
        ng = 64 // gs
        view = (256, ng, gs * 64 * 64)
        st_ = torch.zeros(ng, device="cuda")
        th_ = torch.ones(ng, device="cuda")
        bench("  meanabs only, grouped gs=%d" % gs, lambda i: K.meanabs(xs[i], st_, view), 4 * n)
        bench("  absmax only, grouped gs=%d" % gs, lambda i: K.absmax(xs[i], st_, view), 4 * n)
        bench("  sweep only (where_le), grouped gs=%d" % gs,
              lambda i: K.qdq(xs[i], ys[i], th_, 255, _lib.CLIP_WHERE_LE, "write", view=view), 8 * n)
    st1 = torch.zeros(1, device="cuda")
    bench("  meanabs only, whole tensor", lambda i: K.meanabs(xs[i], st1), 4 * n)
    bench("  absmax only, whole tensor", lambda i: K.absmax(xs[i], st1), 4 * n)

    # ---- first tier on the same tensor ----
    o = op("Quantization_int8_V2", quant_mode="minmax", is_weight=False)
    aux = torch.ones(1, device="cuda")
    bench("V2 act fwd (reduce max + sweep)", lambda i: o.forward(True, ["write"], [xs[i]], [ys[i]], [aux]), 12 * n)
    bench("V2 act bwd (STE copy)",
          lambda i: o.backward(["write"], [dys[i]], [xs[i]], [ys[i]], [ys[(i + 1) % NB]], [aux]), 8 * n)
    o = op("ClipGrad_Quantization_int8", quant_mode="minmax", is_weight=False)
    o.init = False
    bench("ClipGrad act fwd (reduce max + clip sweep)", lambda i: o.forward(True, ["write"], [xs[i]], [ys[i]], [aux]), 12 * n)
    bench("ClipGrad act bwd (open-interval mask)",
          lambda i: o.backward(["write"], [dys[i]], [xs[i]], [ys[i]], [ys[(i + 1) % NB]], [aux]), 12 * n)
    auxd = torch.ones(1, device="cuda")
    bench("fold-BN data fwd (2*mean + clip/scale sweep)", lambda i: K.foldbn_data_fwd(xs[i], ys[i], auxd, False, 0.99), 12 * n)

    # ---- second tier ----
    gamma = torch.tensor([0.8], device="cuda")
    dgamma = torch.zeros(1, device="cuda")
    o = op("PACT_PY", nbits=4)
    bench("PACT fwd (sweep)", lambda i: o.forward(True, ["write"], [xs[i], gamma], [ys[i]], []), 8 * n)
    bench("PACT bwd (mask + sum)",
          lambda i: o.backward(["write", "write"], [dys[i]], [xs[i], gamma], [ys[i]], [ys[(i + 1) % NB], dgamma], []), 12 * n)
    o = op("QUANT_STE_PY", nbits=8)
    bench("QUANT_STE fwd (reduce max + sweep)", lambda i: o.forward(True, ["write"], [xs[i]], [ys[i]], []), 12 * n)
    o = op("DoReFa_PY", nbits=4)
    bench("DoReFa fwd (max + tanh sweep)", lambda i: o.forward(True, ["write"], [xs[i]], [ys[i]], []), 12 * n)
    bench("DoReFa bwd (sum + sweep)", lambda i: o.backward(["write"], [dys[i]], [xs[i]], [ys[i]], [ys[(i + 1) % NB]], []), 20 * n)
    pp, cp, gm = torch.tensor([0.1], device="cuda"), torch.tensor([0.9], device="cuda"), torch.ones(1, device="cuda")
    dpp, dcp, dgm = torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda"), torch.zeros(1, device="cuda")
    o = op("QIL_PY", is_weight=False, fix_gamma=True, nbits=4)
    bench("QIL fwd (sweep)", lambda i: o.forward(True, ["write"], [xs[i], pp, cp, gm], [ys[i]], []), 8 * n)
    bench("QIL bwd (mask + 2 sums)", lambda i: o.backward(["write"] * 4, [dys[i]], [xs[i], pp, cp, gm], [ys[i]],
                                                          [ys[(i + 1) % NB], dpp, dcp, dgm], []), 12 * n)
    o = op("WNQ_PY", nbits=4, is_perchannel=False)
    bench("WNQ fwd per-tensor (reduce max + sweep)", lambda i: o.forward(True, ["write"], [xs[i]], [ys[i]], []), 12 * n)
    bench("WNQ bwd per-tensor (max + sum + sweep)", lambda i: o.backward(["write"], [dys[i]], [xs[i]], [ys[i]], [ys[(i + 1) % NB]], []), 24 * n)
    o = op("CLIP_RELU_PY", nbits=8, threshold=8.0)
    bench("CLIP_RELU fwd (sweep)", lambda i: o.forward(True, ["write"], [xs[i]], [ys[i]], []), 8 * n)
    thr1 = torch.ones(1, device="cuda")
    bench("int8 export (codes + step)", lambda i: K.export_int8(xs[i], thr1, 127, _lib.CLIP_SYM), 5 * n)
    # ---- batch statistics of the conv output + fold + per-channel weight QDQ + bias, one launch ----
    for cshape, wshape in (((256, 64, 64, 64), (64, 64, 3, 3)), ((256, 1024, 16, 16), (1024, 512, 1, 1))):
        c = cshape[1]
        conv = [x_.view(cshape) for x_ in xs]
        w_ = torch.randn(wshape, device="cuda") * 0.05
        wq_, bias_, aw_ = torch.empty_like(w_), torch.empty(c, device="cuda"), torch.ones(c, device="cuda")
        mu_, var_ = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
        gmm_, bt_ = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")
        bench("BN batch stats + fold + weight QDQ + bias, conv out %s (one launch)" % "x".join(map(str, cshape)),
              lambda i: K.bnstat_foldbn_weight_fwd(conv[i], mu_, var_, w_, wq_, bias_, aw_, gmm_, bt_, 1e-5, True, True, True),
              4 * n + 8 * w_.numel())
    del xs, ys, dys
    torch.cuda.empty_cache()

    # ---- weights: per-channel and fold-BN (launch-bound: read the microseconds) ----
    for wshape in ((512, 512, 3, 3), (1024, 1, 3, 3), (1000, 2048), (2048, 1024, 1, 1)):
        ws = [torch.randn(wshape, device="cuda") * 0.05 for _ in range(NB)]
        wq = [torch.empty(wshape, device="cuda") for _ in range(NB)]
        m = ws[0].numel()
        for pc in (False, True):
            o = op("Quantization_int8_V2", quant_mode="minmax", is_weight=True, is_weight_perchannel=pc)
            a = torch.ones(wshape[0] if pc else 1, device="cuda")
            bench("V2 weight %s per_channel=%s" % ("x".join(map(str, wshape)), pc),
                  lambda i: o.forward(True, ["write"], [ws[i]], [wq[i]], [a]), 12 * m, reps=48)
        if len(wshape) == 4:
            c = wshape[0]
            gmm, bt, mu, var = (torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda"),
                                torch.randn(c, device="cuda"), torch.rand(c, device="cuda") + 0.5)
            bias, aw = torch.empty(c, device="cuda"), torch.ones(c, device="cuda")
            bench("fold-BN weight %s per-channel" % "x".join(map(str, wshape)),
                  lambda i: K.foldbn_weight_fwd(ws[i], wq[i], bias, aw, gmm, bt, mu, var, 1e-5, True, True, True),
                  12 * m, reps=48)
    return rows


if __name__ == "__main__":
    import torch
    from b200quant import _lib
    torch.cuda.set_device(0)
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    out = run(torch, _lib.context(0), peak, quick=False, verbose=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "microbench2.json"), "w"), indent=1)
