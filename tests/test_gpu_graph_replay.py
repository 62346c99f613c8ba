"""CUDA-graph safety: a captured step must keep producing correct thresholds when it is REPLAYED on new data, in
particular when the new maximum is SMALLER than the previous one (stale statistics would survive an atomicMax whose
tag never changes).  Covers the fused deferred forward, the multi-tensor weight launches and the peer exchange."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _graph(torch, fn):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    return g


@pytest.mark.parametrize("op_type", ["Quantization_int8_V2", "ClipGrad_Quantization_int8"])
@pytest.mark.parametrize("peer", [False, True])
def test_activation_forward_graph_replay_tracks_new_data(op_type, peer):
    import torch
    import b200quant
    from b200quant.dist import attach_peer_exchange
    mk = lambda: b200quant.get_prop(op_type)(quant_mode="minmax", is_weight="False").create_operator(None, None, None)
    cap, ref = mk(), mk()
    ex = attach_peer_exchange([cap], torch.device("cuda", 0)) if peer else None
    shape = (8, 64, 28, 28)
    x = torch.randn(shape, device="cuda") * 10
    y, aux = torch.zeros_like(x), torch.ones(1, device="cuda")
    yr, auxr = torch.zeros_like(x), torch.ones(1, device="cuda")
    if op_type == "ClipGrad_Quantization_int8":   # leave the first-batch state before capturing (steady state)
        cap.forward(True, ["write"], [x], [y], [aux])
        ref.forward(True, ["write"], [x], [yr], [auxr])
    g = _graph(torch, lambda: cap.forward(True, ["write"], [x], [y], [aux]))
    aux.copy_(auxr)                                # capture ran the step twice: restart both from the same state
    gen = torch.Generator(device="cuda").manual_seed(5)
    for scale in (5.0, 0.3, 0.01, 2.0, 0.001):     # maxima going DOWN must be seen
        x.copy_(torch.randn(shape, device="cuda", generator=gen) * scale)
        g.replay()
        ref.forward(True, ["write"], [x], [yr], [auxr])
        torch.cuda.synchronize()
        assert torch.equal(aux.view(torch.int32), auxr.view(torch.int32)), scale
        assert torch.equal(y.view(torch.int32), yr.view(torch.int32)), scale
    if ex is not None:
        ex.close()


def test_weight_group_graph_replay_tracks_new_data():
    import torch
    import b200quant
    from b200quant.multi import WeightGroup
    shapes = [((64, 3, 7, 7), False), ((256, 64, 1, 1), False), ((32, 1, 3, 3), True), ((16, 600), True)]
    mk = lambda pc: b200quant.get_prop("Quantization_int8_V2")(quant_mode="minmax", is_weight="True",
                                                               is_weight_perchannel=str(pc)).create_operator(None, None, None)
    ops, refs = [mk(pc) for _, pc in shapes], [mk(pc) for _, pc in shapes]
    ws = [torch.randn(s, device="cuda") for s, _ in shapes]
    ys, yrs = [torch.zeros_like(w) for w in ws], [torch.zeros_like(w) for w in ws]
    auxs = [torch.ones(s[0] if pc else 1, device="cuda") for s, pc in shapes]
    auxrs = [a.clone() for a in auxs]
    group = WeightGroup(ops, ws, ys, auxs)
    g = _graph(torch, lambda: group.forward(True))
    for scale in (3.0, 0.2, 0.004, 1.0):
        for w in ws:
            w.copy_(torch.randn_like(w) * scale)
        g.replay()
        for op, w, y, a in zip(refs, ws, yrs, auxrs):
            op.forward(True, ["write"], [w], [y], [a])
        torch.cuda.synchronize()
        for i in range(len(ws)):
            assert torch.equal(auxs[i].view(torch.int32), auxrs[i].view(torch.int32)), (scale, i)
            assert torch.equal(ys[i].view(torch.int32), yrs[i].view(torch.int32)), (scale, i)
    group.close()


def test_tuning_options_never_change_results():
    """`pdl` (programmatic dependent launch), `deferred`, `reverse`, `fast_div` and the grid knobs are performance options:
    a forward + backward chain of several operators must give the same bits whatever they are set to."""
    import torch
    import b200quant
    from b200quant import _lib
    ctx = _lib.context(0)
    gen = torch.Generator(device="cuda").manual_seed(17)
    xs = [torch.randn(s, device="cuda", generator=gen) * 3 for s in [(8, 64, 28, 28), (4, 16, 7, 7), (3, 5, 11)]]
    dys = [torch.randn_like(x) for x in xs]
    w = torch.randn(64, 32, 3, 3, device="cuda", generator=gen) * 0.05

    def chain():
        outs = []
        for op_type, attrs in (("Quantization_int8_V2", dict(quant_mode="minmax", is_weight="False")),
                               ("ClipGrad_Quantization_int8", dict(quant_mode="minmax", is_weight="False")),
                               ("GDRQ_PY", dict(nbits="8", group_size="-1", is_weight="False")),
                               ("GDRQ_PY", dict(nbits="4", group_size="4", is_weight="False"))):
            for x, dy in zip(xs, dys):
                if op_type == "GDRQ_PY" and attrs["group_size"] != "-1" and (x.dim() < 2 or x.shape[1] % 4):
                    continue
                op = b200quant.get_prop(op_type)(**attrs).create_operator(None, None, None)
                groups = 1 if attrs.get("group_size", "-1") == "-1" else x.shape[1] // 4
                aux = torch.ones(groups, device="cuda")
                y, dx = torch.zeros_like(x), torch.zeros_like(x)
                op.forward(True, ["write"], [x], [y], [aux])
                op.backward(["write"], [dy], [x], [y], [dx], [aux])
                outs += [y, dx, aux]
        opw = b200quant.get_prop("Quantization_int8_V2")(quant_mode="minmax", is_weight="True",
                                                         is_weight_perchannel="True").create_operator(None, None, None)
        wq, aw = torch.zeros_like(w), torch.ones(64, device="cuda")
        opw.forward(True, ["write"], [w], [wq], [aw])
        outs += [wq, aw]
        torch.cuda.synchronize()
        return [o.clone() for o in outs]

    base = chain()
    defaults = {k: ctx.get_option(k) for k in ("pdl", "deferred", "reverse", "fast_div", "blocks_per_sm",
                                                "reduce_blocks_per_sm")}
    try:
        for key, value in (("pdl", 0), ("deferred", 0), ("reverse", 0), ("fast_div", 0), ("blocks_per_sm", 2),
                           ("reduce_blocks_per_sm", 1)):
            ctx.set_option(key, value)
            got = chain()
            ctx.set_option(key, defaults[key])
            for a, b in zip(base, got):
                assert torch.equal(a.view(torch.int32), b.view(torch.int32)), key
    finally:
        for k, v in defaults.items():
            ctx.set_option(k, v)
