"""Per-kernel micro-benchmark on one GPU (CUDA events, inputs larger than L2 or L2 flushed between runs).
Writes gpurun_out/microbench.json.  Usage: python tools/microbench.py [--quick]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import b200quant  # noqa: E402,F401
from b200quant import _kernels as K  # noqa: E402
from b200quant import _lib  # noqa: E402


def timeit(fn, iters, flush=None):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    for a, b in ev:
        if flush is not None:
            flush.add_(1.0)
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], ts[0]


def main():
    quick = "--quick" in sys.argv
    torch.cuda.set_device(0)
    ctx = _lib.context(0)
    flush = torch.zeros(64 * 1024 * 1024, device="cuda")  # 256 MiB > L2
    sizes = [1 << 20, 1 << 24, 1 << 26, 1 << 28, 256 * 64 * 112 * 112]
    if quick:
        sizes = [1 << 24, 256 * 64 * 112 * 112]
    out = {"device": torch.cuda.get_device_name(0), "sms": ctx.num_sms(), "rows": []}
    for n in sizes:
        x = torch.empty(n, device="cuda").uniform_(-1, 1)
        y = torch.empty_like(x)
        dy = torch.randn(n, device="cuda")
        dx = torch.empty_like(x)
        aux = torch.ones(1, device="cuda")
        stat = torch.zeros(1, device="cuda")
        thr = torch.tensor([1.0], device="cuda")
        small = n * 4 < 200e6
        fl = flush if small else None
        iters = 20 if n <= (1 << 26) else 10
        rows = []

        def add(name, fn, bytes_per_elem):
            med, best = timeit(fn, iters, fl)
            rows.append(dict(kernel=name, n=n, ms_median=med, ms_best=best, alg_bytes=bytes_per_elem * n,
                             gbs_median=bytes_per_elem * n / med / 1e6, gbs_best=bytes_per_elem * n / best / 1e6))
            print("%-34s n=%-11d %8.3f ms  %8.1f GB/s (best %8.1f)" % (name, n, med, rows[-1]["gbs_median"],
                                                                     rows[-1]["gbs_best"]), flush=True)

        add("absmax (reduce_flat max)", lambda: K.absmax(x, stat), 4)
        add("meanabs (reduce_flat sum)", lambda: K.meanabs(x, stat), 4)
        add("qdq none (hot)", lambda: K.qdq(x, y, thr, 127, _lib.CLIP_NONE, "write"), 8)
        add("qdq clip_sym (hot)", lambda: K.qdq(x, y, thr, 127, _lib.CLIP_SYM, "write"), 8)
        ctx.set_option("fast_div", 0)
        add("qdq clip_sym exact-div", lambda: K.qdq(x, y, thr, 127, _lib.CLIP_SYM, "write"), 8)
        ctx.set_option("fast_div", 1)
        add("ste bwd copy", lambda: K.ste_bwd(dy, dx, "write"), 8)
        add("clipgrad bwd mask", lambda: K.clipgrad_bwd(x, dy, dx, aux), 12)
        add("torch copy_ (reference peak)", lambda: y.copy_(x), 8)
        add("V2 act fwd fused (reduce+qdq)",
            lambda: K.minmax_quant_fwd(0, x, y, aux, False, False, True, False, 0.99, "write"), 12)
        add("ClipGrad act fwd fused", lambda: K.minmax_quant_fwd(1, x, y, aux, False, False, True, False, 0.99, "write"), 12)
        for rev in (0, 1):
            ctx.set_option("reverse", rev)
            add("V2 act fwd fused reverse=%d" % rev,
                lambda: K.minmax_quant_fwd(0, x, y, aux, False, False, True, False, 0.99, "write"), 12)
        ctx.set_option("reverse", 1)
        for bps in (2, 4, 6, 8, 12, 16):
            ctx.set_option("blocks_per_sm", bps)
            add("qdq clip_sym bps=%d" % bps, lambda: K.qdq(x, y, thr, 127, _lib.CLIP_SYM, "write"), 8)
            add("absmax bps=%d" % bps, lambda: K.absmax(x, stat), 4)
        ctx.set_option("blocks_per_sm", 8)
        out["rows"].extend(rows)
        del x, y, dy, dx
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "microbench.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
