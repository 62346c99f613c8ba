"""Launches tests/multi_gpu_check.py under torchrun when at least two GPUs are visible (skipped on 1-GPU boxes).
On one GPU the peer-memory kernels are still exercised with world = 1 (the sweep polls a mailbox the preceding
reduction of the same stream already filled, so nothing waits on another launch)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_peer_exchange_world1_matches_plain_path():
    import torch
    import b200quant
    from b200quant.dist import attach_peer_exchange
    for op_type in ("Quantization_int8_V2", "ClipGrad_Quantization_int8"):
        mk = lambda: b200quant.get_prop(op_type)(quant_mode="minmax", is_weight="False").create_operator(None, None, None)
        a, b = mk(), mk()
        ex = attach_peer_exchange([b], torch.device("cuda", 0))
        aux_a, aux_b = torch.ones(1, device="cuda"), torch.ones(1, device="cuda")
        g = torch.Generator(device="cuda").manual_seed(5)
        for step, shape in enumerate([(8, 32, 14, 14), (5, 3, 7), (2, 64, 56, 56), (1, 9)]):
            x = torch.randn(shape, device="cuda", generator=g) * (1 + step)
            ya, yb = torch.zeros_like(x), torch.zeros_like(x)
            a.forward(True, ["write"], [x], [ya], [aux_a])
            b.forward(True, ["write"], [x], [yb], [aux_b])
            torch.cuda.synchronize()
            assert torch.equal(aux_a.view(torch.int32), aux_b.view(torch.int32)), (op_type, step)
            assert torch.equal(ya.view(torch.int32), yb.view(torch.int32)), (op_type, step)
        ex.close()


@pytest.mark.parametrize("peer_mode", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("reverse_all", [False, True])
def test_peer_sweep_variants_equal_plain_path_around_tile_edges(peer_mode, reverse_all):
    """peer_mode 2 / 3 give every block 2 / 3 tiles (one in registers, the others staged in shared memory by bulk
    copies issued while the block waits for the peers' statistic): tensor sizes around the 4096-element tile and the
    super-tile boundaries, odd heads and tails, ascending and descending tile order, both staging points."""
    import torch
    import b200quant
    from b200quant import _lib
    from b200quant.dist import attach_peer_exchange
    ctx = _lib.context(0)
    saved = {k: ctx.get_option(k) for k in ("peer_mode", "reverse_min_mb", "peer_stage_early")}
    g = torch.Generator(device="cuda").manual_seed(11)
    sizes = [1, 7, 8, 4095, 4096, 4097, 8191, 8192, 8200, 12287, 12288, 12296, 3 * 4096 + 8, 5 * 4096 - 8, 6 * 4096,
             7 * 4096 + 24, 100003, 1 << 20, (1 << 22) + 40]
    try:
        for early in (0, 1):
            ctx.set_option("peer_mode", peer_mode)
            ctx.set_option("peer_stage_early", early)
            if reverse_all:
                ctx.set_option("reverse_min_mb", 0)
            for op_type in ("Quantization_int8_V2", "ClipGrad_Quantization_int8"):
                mk = lambda: b200quant.get_prop(op_type)(quant_mode="minmax", is_weight="False").create_operator(None, None, None)
                a, b = mk(), mk()
                ex = attach_peer_exchange([b], torch.device("cuda", 0))
                aux_a, aux_b = torch.ones(1, device="cuda"), torch.ones(1, device="cuda")
                for step, n in enumerate(sizes):
                    for off in (0, 3):                      # a misaligned view: scalar head and tail
                        buf = torch.randn(n + 8, device="cuda", generator=g) * (1 + step % 5)
                        x = buf[off:off + n]
                        ybuf_a, ybuf_b = torch.full_like(buf, 7.0), torch.full_like(buf, 7.0)
                        a.forward(True, ["write"], [x], [ybuf_a[off:off + n]], [aux_a])
                        b.forward(True, ["write"], [x], [ybuf_b[off:off + n]], [aux_b])
                        assert torch.equal(aux_a.view(torch.int32), aux_b.view(torch.int32)), (op_type, n, off)
                        assert torch.equal(ybuf_a.view(torch.int32), ybuf_b.view(torch.int32)), (op_type, n, off)
                torch.cuda.synchronize()
                ex.close()
            if peer_mode < 2:
                break
    finally:
        for k, v in saved.items():
            ctx.set_option(k, v)


def test_peer_exchange_world1_mean_based_ops():
    """Whole-tensor GDRQ_PY activations and the GDRQ_Fold_BN data path through b2q_peer_meanabs_quant_fwd_f32."""
    import torch
    import b200quant
    from b200quant.dist import attach_peer_exchange
    g = torch.Generator(device="cuda").manual_seed(21)
    mk = lambda: b200quant.get_prop("GDRQ_PY")(nbits="8", group_size="-1", is_weight="False",
                                               delay_quant="1").create_operator(None, None, None)
    a, b = mk(), mk()
    mkf = lambda: b200quant.get_prop("GDRQ_Fold_BN")(
        quant_mode="minmax", is_weight_perchannel="False", name="c", num_filter="8", num_group="1", kernel="(1, 1)",
        stride="(1, 1)", pad="(0, 0)", no_bias="True").create_operator(None, None, None)
    fa, fb = mkf(), mkf()
    ex = attach_peer_exchange([b, fb], torch.device("cuda", 0))
    assert b.peer is ex and fb.peer is ex
    al_a, al_b = torch.ones(1, device="cuda"), torch.ones(1, device="cuda")
    for step, shape in enumerate([(4, 16, 9, 9), (2, 16, 56, 56), (3, 5, 7), (64, 64, 28, 28)]):
        x = torch.randn(shape, device="cuda", generator=g) * (1 + step)
        ya, yb = torch.zeros_like(x), torch.zeros_like(x)
        a.forward(True, ["write"], [x], [ya], [al_a])
        b.forward(True, ["write"], [x], [yb], [al_b])
        assert torch.equal(al_a.view(torch.int32), al_b.view(torch.int32)), step
        assert torch.equal(ya.view(torch.int32), yb.view(torch.int32)), step
    w = torch.randn(8, 4, 1, 1, device="cuda", generator=g) * 0.2
    gamma, beta = torch.rand(8, device="cuda", generator=g) + 0.5, torch.randn(8, device="cuda", generator=g)
    mean, var = torch.randn(8, device="cuda", generator=g), torch.rand(8, device="cuda", generator=g) + 0.1
    aux_a = [torch.ones(1, device="cuda"), torch.ones(1, device="cuda")]
    aux_b = [torch.ones(1, device="cuda"), torch.ones(1, device="cuda")]
    for step in range(3):
        x = torch.randn(2, 4, 30, 30, device="cuda", generator=g) * (1 + step)
        bn_out = torch.zeros(2, 8, 30, 30, device="cuda")
        ya, yb = torch.zeros_like(bn_out), torch.zeros_like(bn_out)
        fa.forward(True, ["write"], [x, w, bn_out, gamma, beta, mean, var], [ya], aux_a)
        fb.forward(True, ["write"], [x, w, bn_out, gamma, beta, mean, var], [yb], aux_b)
        assert torch.equal(aux_a[0].view(torch.int32), aux_b[0].view(torch.int32)), step
        assert torch.equal(fa.data_q.view(torch.int32), fb.data_q.view(torch.int32)), step
    torch.cuda.synchronize()
    ex.close()


def test_mean_based_ops_sync_path_matches_fused_path():
    """GDRQ_PY / GDRQ_Fold_BN with a (world = 1, no-op) ThresholdSync take the reduce -> exchange -> update -> sweep
    route built from the primitives; it has to give the bits of the fused entry points."""
    import torch
    import b200quant
    from b200quant.dist import ThresholdSync
    g = torch.Generator(device="cuda").manual_seed(9)
    for group_size in (-1, 2, 8):
        mk = lambda: b200quant.get_prop("GDRQ_PY")(nbits="4", group_size=str(group_size), is_weight="False",
                                                   delay_quant="1").create_operator(None, None, None)
        a, b = mk(), mk()
        b.sync = ThresholdSync()
        groups = 1 if group_size == -1 else 16 // group_size
        al_a, al_b = torch.ones(groups, device="cuda"), torch.ones(groups, device="cuda")
        for step, shape in enumerate([(4, 16, 9, 9), (2, 16, 28, 28), (3, 16, 1, 5)]):
            x = torch.randn(shape, device="cuda", generator=g) * (1 + step)
            ya, yb = torch.zeros_like(x), torch.zeros_like(x)
            a.forward(True, ["write"], [x], [ya], [al_a])
            b.forward(True, ["write"], [x], [yb], [al_b])
            assert torch.equal(al_a.view(torch.int32), al_b.view(torch.int32)), (group_size, step)
            assert torch.equal(ya.view(torch.int32), yb.view(torch.int32)), (group_size, step)
        assert b.sync.calls == 3

    mk = lambda: b200quant.get_prop("GDRQ_Fold_BN")(
        quant_mode="minmax", is_weight_perchannel="True", name="c", num_filter="8", num_group="1", kernel="(3, 3)",
        stride="(1, 1)", pad="(1, 1)", no_bias="True").create_operator(None, None, None)
    a, b = mk(), mk()
    b.sync = ThresholdSync()
    w = torch.randn(8, 4, 3, 3, device="cuda", generator=g) * 0.2
    gamma, beta = torch.rand(8, device="cuda", generator=g) + 0.5, torch.randn(8, device="cuda", generator=g)
    mean, var = torch.randn(8, device="cuda", generator=g), torch.rand(8, device="cuda", generator=g) + 0.1
    aux_a = [torch.ones(1, device="cuda"), torch.ones(8, device="cuda")]
    aux_b = [torch.ones(1, device="cuda"), torch.ones(8, device="cuda")]
    for step in range(3):
        x = torch.randn(2, 4, 10, 10, device="cuda", generator=g) * (1 + step)
        bn_out = torch.zeros(2, 8, 10, 10, device="cuda")
        ya, yb = torch.zeros_like(bn_out), torch.zeros_like(bn_out)
        a.forward(True, ["write"], [x, w, bn_out, gamma, beta, mean, var], [ya], aux_a)
        b.forward(True, ["write"], [x, w, bn_out, gamma, beta, mean, var], [yb], aux_b)
        assert torch.equal(aux_a[0].view(torch.int32), aux_b[0].view(torch.int32)), step
        assert torch.equal(a.data_q.view(torch.int32), b.data_q.view(torch.int32)), step
        assert torch.equal(ya.view(torch.int32), yb.view(torch.int32)), step


def test_peer_allreduce_and_grouped_statistics_at_world1():
    """b2q_peer_allreduce_{sum,max}_f32 and the grouped GDRQ_PY exchange through them on one GPU (world 1: the flag
    barriers go through the own mailbox): sum / max are the identity, average too, nothing outside `count` is touched;
    grouped GDRQ activations with the exchange attached give the bits of the plain path."""
    import torch
    import b200quant
    from b200quant.dist import PeerBuffer, PeerGradBucket, PeerThresholdExchange, attach_peer_exchange
    dev = torch.device("cuda", 0)
    ex = PeerThresholdExchange(dev)
    g = torch.Generator(device="cuda").manual_seed(2)
    buf = PeerBuffer(ex, 100003)
    for count in (1, 3, 4, 5, 4099, 100003):
        for kind in ("sum", "avg", "max"):
            x = torch.randn(count, device="cuda", generator=g)
            buf.tensor.fill_(7.0)
            buf.tensor[:count].copy_(x)
            if kind == "max":
                buf.allreduce_max(count)
            else:
                buf.allreduce_sum(count, average=(kind == "avg"))
            assert torch.equal(buf.tensor[:count].view(torch.int32), x.view(torch.int32)), (count, kind)
            assert bool((buf.tensor[count:] == 7.0).all())
    bucket = PeerGradBucket([(8, 3, 3, 3), (10, 8), (5,)], ex)
    assert [tuple(v.shape) for v in bucket.views] == [(8, 3, 3, 3), (10, 8), (5,)] and bucket.flat.numel() == 216 + 80 + 5
    bucket.views[1].fill_(2.0)
    bucket.allreduce(average=True)
    assert float(bucket.flat.sum()) == 160.0
    torch.cuda.synchronize()
    assert ex.status() is None
    bucket.close()
    buf.close()
    ex.close()
    mk = lambda: b200quant.get_prop("GDRQ_PY")(nbits="8", group_size="4", is_weight="False").create_operator(None, None, None)
    a, b = mk(), mk()
    ex = attach_peer_exchange([b], dev)
    assert b.peer is ex and b.sync is None
    al_a, al_b = torch.ones(4, device="cuda"), torch.ones(4, device="cuda")
    for step, shape in enumerate([(4, 16, 9, 9), (2, 16, 56, 56), (3, 16, 7, 5)]):
        x = torch.randn(shape, device="cuda", generator=g) * (1 + step)
        ya, yb = torch.zeros_like(x), torch.zeros_like(x)
        a.forward(True, ["write"], [x], [ya], [al_a])
        b.forward(True, ["write"], [x], [yb], [al_b])
        assert torch.equal(al_a.view(torch.int32), al_b.view(torch.int32)), step
        assert torch.equal(ya.view(torch.int32), yb.view(torch.int32)), step
    torch.cuda.synchronize()
    ex.close()


def test_two_rank_threshold_exchange():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "failures=0" in res.stdout
