"""A/B of the segmented reductions' tuning options on 2^26-element tensors: the batch-statistics + fold kernel and the
grouped mean|x| / max|x| reductions, register-staged (stream_reduce=0) vs the TMA-staged ring (stages, integer-pipe
conversions, pieces per SM)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200quant  # noqa: E402,F401
from b200quant import _kernels as K  # noqa: E402
from b200quant import _lib  # noqa: E402

torch.cuda.set_device(0)
ctx = _lib.context(0)
peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6545.6
NB = 6
n = 1 << 26
xs = [torch.empty(n, device="cuda").uniform_(-1, 1) for _ in range(NB)]


def timeit(fn, reps=12):
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


CONFIGS = [dict(seg_masked=0), dict(seg_masked=1)]

for cshape, wshape in (((256, 64, 64, 64), (64, 64, 3, 3)), ((256, 1024, 16, 16), (1024, 512, 1, 1)),
                       ((128, 512, 28, 28), (512, 128, 1, 1)), ((256, 1024, 14, 14), (1024, 256, 1, 1)),
                       ((256, 2048, 7, 7), (2048, 512, 1, 1)), ((256, 1024, 7, 7), (1024, 1, 3, 3))):
    c = cshape[1]
    m = int(torch.Size(cshape).numel())
    conv = [x[:m].view(cshape) for x in xs]
    w = torch.randn(wshape, device="cuda") * 0.05
    wq, bias, aw = torch.empty_like(w), torch.empty(c, device="cuda"), torch.ones(c, device="cuda")
    mu, var = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    gm, bt = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")
    for cfg in CONFIGS:
        for k, v in cfg.items():
            ctx.set_option(k, v)
        ms = timeit(lambda i: K.bnstat_foldbn_weight_fwd(conv[i], mu, var, w, wq, bias, aw, gm, bt, 1e-5, True, True, True))
        gbs = 4 * m / ms / 1e6
        print("bnstat %-16s %-70s %7.1f us %7.1f GB/s %.3f" % ("x".join(map(str, cshape)), cfg, ms * 1e3, gbs, gbs / peak),
              flush=True)

