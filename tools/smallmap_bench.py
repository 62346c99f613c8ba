"""Grouped GDRQ_PY activations (one alpha per channel group) on feature maps whose rows are not a multiple of eight
floats: forward (per-group mean|x| + alpha update + clip / QDQ sweep) and backward (|x| <= alpha mask), microseconds and
fraction of the measured copy peak on the algorithmic bytes (12 B/element each)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200quant  # noqa: E402

torch.cuda.set_device(0)
peak = 6545.6
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
NB = 4
for shape, gs in (((256, 256, 28, 28), 1), ((256, 1024, 14, 14), 1), ((256, 1024, 14, 14), 16), ((256, 2048, 7, 7), 1),
                  ((256, 2048, 7, 7), 32)):
    n = int(torch.Size(shape).numel())
    xs = [torch.empty(shape, device="cuda").uniform_(-1, 1) for _ in range(NB)]
    ys = [torch.empty(shape, device="cuda") for _ in range(NB)]
    dys = [torch.empty(shape, device="cuda").normal_() for _ in range(NB)]
    op = b200quant.get_prop("GDRQ_PY")(nbits="8", group_size=str(gs), is_weight="False").create_operator(None, None, None)
    alpha = torch.ones(shape[1] // gs, device="cuda")

    def timeit(fn, reps=10):
        for i in range(3):
            fn(i % NB)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            fn(i % NB)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    f = timeit(lambda i: op.forward(True, ["write"], [xs[i]], [ys[i]], [alpha]))
    b = timeit(lambda i: op.backward(["write"], [dys[i]], [xs[i]], [ys[i]], [ys[(i + 1) % NB]], [alpha]))
    print("GDRQ act %-16s group_size=%-2d  fwd %7.1f us %.3f   bwd %7.1f us %.3f" % (
        "x".join(map(str, shape)), gs, f * 1e3, 12 * n / f / 1e6 / peak, b * 1e3, 12 * n / b / 1e6 / peak), flush=True)
