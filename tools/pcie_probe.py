"""What does the host link deliver?  Pinned H2D alone, D2H alone, and both directions at once (two streams)."""
import json
import os
import sys
import time

import torch

torch.cuda.set_device(0)
n = 256 * 1024 * 1024  # 1 GiB of float32
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_a = torch.empty(n, device="cuda")
d_b = torch.empty(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
res = {}


def run(name, fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    res[name] = 4 * n / dt / 1e9
    print("%-28s %7.1f GB/s per direction" % (name, res[name]), flush=True)


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


run("h2d only", h2d)
run("d2h only", d2h)
run("h2d + d2h concurrently", both)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/pcie_probe.json", "w"))

# ---- the library's host-buffer pipeline on the same link ----
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import b200quant  # noqa: E402,F401
from b200quant import _kernels as K  # noqa: E402
from b200quant import _lib  # noqa: E402

ctx = _lib.context(0)
for mb in (64, 256, 822):
    m = mb * 1024 * 1024 // 4
    xs = [torch.empty(m, dtype=torch.float32).uniform_(-1, 1).pin_memory() for _ in range(2)]
    ys = [torch.empty(m, dtype=torch.float32).pin_memory() for _ in range(2)]
    auxs = [torch.ones(1).pin_memory() for _ in range(8)]

    def pipe(reps=8):
        for i in range(reps):
            K.minmax_quant_fwd(0, xs[i & 1], ys[i & 1], auxs[i % 8], False, False, True, False, 0.99, "write")
        ctx.host_sync()

    pipe(4)
    t0 = time.perf_counter()
    pipe(8)
    dt = time.perf_counter() - t0
    res["host_pipeline_%dMB" % mb] = 8 * 4 * m / dt / 1e9
    print("library host pipeline, %4d MB tensors: %6.1f GB/s per direction" % (mb, res["host_pipeline_%dMB" % mb]), flush=True)
json.dump(res, open("gpurun_out/pcie_probe.json", "w"))
