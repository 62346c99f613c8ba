"""Per-node cost of the small-tensor forward, two PDL-chained launches vs one cluster launch, by tensor size: 64 nodes of
the same size back to back (distinct tensors), CUDA events, microseconds per node."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import b200quant  # noqa: E402
from b200quant import _lib  # noqa: E402

torch.cuda.set_device(0)
ctx = _lib.context(0)
NODES = 64


def bench(op_type, attrs, n, cfg):
    for k, v in cfg.items():
        ctx.set_option(k, v)
    ops = [b200quant.get_prop(op_type)(**attrs).create_operator(None, None, None) for _ in range(NODES)]
    xs = [torch.randn(n, device="cuda") * 0.1 for _ in range(NODES)]
    ys = [torch.empty(n, device="cuda") for _ in range(NODES)]
    auxs = [torch.ones(1, device="cuda") for _ in range(NODES)]

    def step():
        for o, x, y, a in zip(ops, xs, ys, auxs):
            o.forward(True, ["write"], [x], [y], [a])
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # captured in a graph so that the host launch rate is out of the picture
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        step()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            step()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    graph_us = e0.elapsed_time(e1) * 1e3 / 10 / NODES
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    eager_us = e0.elapsed_time(e1) * 1e3 / 10 / NODES
    return graph_us, eager_us


for op_type, attrs in (("Quantization_int8_V2", dict(quant_mode="minmax", is_weight="True")),
                       ("GDRQ_PY", dict(nbits="8", group_size="-1", is_weight="True"))):
    for n in (4096, 9408, 16384, 36864, 65536, 147456, 262144, 327680):
        row = []
        for cfg in (dict(cluster_fwd=0), dict(cluster_fwd=1, cluster_words_per_cta=256), dict(cluster_fwd=1, cluster_words_per_cta=1024),
                    dict(cluster_fwd=1, cluster_words_per_cta=5120)):
            row.append(bench(op_type, attrs, n, cfg))
        print("%-22s n=%7d  two launches %5.2f / %5.2f us   cluster(256 w/CTA) %5.2f / %5.2f   cluster(1024) %5.2f / %5.2f   "
              "cluster(5120) %5.2f / %5.2f   [graph / eager per node]" % ((op_type, n) + tuple(v for r in row for v in r)), flush=True)
