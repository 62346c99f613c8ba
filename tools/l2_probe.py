"""Does the QDQ sweep find what the reduction just read in L2?  For each tensor size: time the QDQ kernel alone
(CUDA events) right after a reduction over (A) the same tensor and (B) a different tensor, with the sweep walking
descending (reverse=1) or ascending (reverse=0) addresses.  Writes gpurun_out/l2_probe.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import b200quant  # noqa: E402,F401
from b200quant import _kernels as K  # noqa: E402
from b200quant import _lib  # noqa: E402

torch.cuda.set_device(0)
ctx = _lib.context(0)
flush = torch.zeros(128 * 1024 * 1024, device="cuda")  # 512 MiB
stat = torch.zeros(1, device="cuda")
thr = torch.tensor([1.0], device="cuda")
rows = []
for mb in (4, 8, 16, 24, 32, 48, 64, 96, 128, 192, 256, 512):
    n = mb * 1024 * 1024 // 4
    x1 = torch.empty(n, device="cuda").uniform_(-1, 1)
    x2 = torch.empty(n, device="cuda").uniform_(-1, 1)
    y = torch.empty(n, device="cuda")
    for rev in (1, 0):
        ctx.set_option("reverse", rev)
        for same in (True, False):
            ts = []
            for it in range(12):
                flush.add_(1.0)
                K.absmax(x1, stat)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                K.qdq(x1 if same else x2, y, thr, 127, _lib.CLIP_NONE, "write")
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            med = ts[len(ts) // 2]
            rows.append(dict(mb=mb, reverse=rev, same_tensor=same, qdq_us=med * 1e3, alg_gbs=8 * n / med / 1e6))
            print("size %4d MiB reverse=%d same=%-5s qdq %8.2f us  %8.1f GB/s (alg)" % (mb, rev, same, med * 1e3,
                                                                                      8 * n / med / 1e6), flush=True)
    del x1, x2, y
ctx.set_option("reverse", 1)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "l2_probe.json"), "w"), indent=1)
