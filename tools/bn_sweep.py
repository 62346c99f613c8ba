"""A/B of the batch-statistics + fold kernel's tuning options (bn_variant, bn_pieces_per_sm) on 2^26-element conv outputs."""
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
for cshape, wshape in (((256, 64, 64, 64), (64, 64, 3, 3)), ((256, 1024, 16, 16), (1024, 512, 1, 1)),
                       ((256, 256, 32, 32), (256, 1, 3, 3))):
    c = cshape[1]
    conv = [x.view(cshape) for x in xs]
    w = torch.randn(wshape, device="cuda") * 0.05
    wq, bias, aw = torch.empty_like(w), torch.empty(c, device="cuda"), torch.ones(c, device="cuda")
    mu, var = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    gm, bt = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")
    for variant in (1, 4, 5, 6):
        for pieces in (4, 8, 12):
            ctx.set_option("bn_variant", variant)
            ctx.set_option("bn_pieces_per_sm", pieces)
            fn = lambda i: K.bnstat_foldbn_weight_fwd(conv[i], mu, var, w, wq, bias, aw, gm, bt, 1e-5, True, True, True)
            for i in range(3):
                fn(i % NB)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(12):
                fn(i % NB)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 12
            gbs = 4 * n / ms / 1e6
            print("conv out %-16s variant %d pieces/SM %2d: %7.1f us %7.1f GB/s %.3f" % (
                "x".join(map(str, cshape)), variant, pieces, ms * 1e3, gbs, gbs / peak), flush=True)
