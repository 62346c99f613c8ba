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


def test_two_rank_threshold_exchange():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "failures=0" in res.stdout
