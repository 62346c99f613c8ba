"""BASELINE.json config 1: symbol/simple.py net with quant_ops int8 fake-quant, CIFAR-shaped 32x32 batch 32, fwd+bwd.
The CUDA path (SimpleCifarNet: our CustomOps under torch autograd) trains next to a CPU twin whose quantization nodes
are the NumPy oracle; both start from the same weights and see the same synthetic batches."""
import numpy as np
import pytest

from oracle import quant_oracle as qo

pytestmark = pytest.mark.gpu
F = np.float32


def _twin(torch, nn, style):
    op_type = {"quant_ops": "Quantization_int8_V2", "int8_api": "ClipGrad_Quantization_int8"}[style]

    class OracleFn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, node, x):
            xn = x.detach().numpy().astype(F)
            y = np.zeros_like(xn)
            node.op.forward(node.training, ["write"], [xn], [y], [node.aux])
            ctx.node, ctx.xn, ctx.y = node, xn, y
            return torch.from_numpy(y)

        @staticmethod
        def backward(ctx, g):
            dx = np.zeros_like(ctx.xn)
            ctx.node.op.backward(["write"], [g.numpy().astype(F)], [ctx.xn], [ctx.y], [dx], [ctx.node.aux])
            return None, torch.from_numpy(dx)

    class Node(nn.Module):
        def __init__(self, is_weight):
            super().__init__()
            self.op = qo.create(op_type, quant_mode="minmax", is_weight=str(is_weight), is_weight_perchannel="False")
            self.aux = np.ones(1, F)

        def forward(self, x):
            return OracleFn.apply(self, x)

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.w1 = nn.Parameter(torch.zeros(8, 3, 3, 3))
            self.w2 = nn.Parameter(torch.zeros(8, 8, 3, 3))
            self.w3 = nn.Parameter(torch.zeros(10, 8))
            self.b3 = nn.Parameter(torch.zeros(10))
            self.bn1 = nn.BatchNorm2d(8, eps=1e-3, momentum=0.1)
            self.bn2 = nn.BatchNorm2d(8, eps=1e-3, momentum=0.1)
            self.q = nn.ModuleList([Node(i % 2 == 0) for i in range(6)])   # w1,d1,w2,d2,w3,d3

        def forward(self, x):
            f = torch.nn.functional
            x = torch.relu(self.bn1(f.conv2d(self.q[1](x), self.q[0](self.w1), None, 2, 1)))
            x = torch.relu(self.bn2(f.conv2d(self.q[3](x), self.q[2](self.w2), None, 2, 1)))
            x = x.mean(dim=(2, 3))
            return f.linear(self.q[5](x), self.q[4](self.w3), self.b3)

    return Net()


@pytest.mark.parametrize("style", ["quant_ops", "int8_api"])
def test_simple_net_training_matches_oracle_twin(style):
    import torch
    import torch.nn as nn
    from b200quant.harness import SimpleCifarNet, export_mx_params
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(5)
    gpu = SimpleCifarNet(style=style).cuda().train()
    cpu = _twin(torch, nn, style).train()
    with torch.no_grad():
        cpu.w1.copy_(gpu.stage1_conv.weight.cpu())
        cpu.w2.copy_(gpu.stage2_conv.weight.cpu())
        cpu.w3.copy_(gpu.fc1.weight.cpu())
        cpu.b3.copy_(gpu.fc1.bias.cpu())
    og = torch.optim.SGD(gpu.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)   # edict_config.py:58-60
    oc = torch.optim.SGD(cpu.parameters(), lr=0.05, momentum=0.9, weight_decay=1e-4)
    rng = np.random.default_rng(5)
    for step in range(4):
        x = torch.from_numpy(rng.uniform(-1, 1, (32, 3, 32, 32)).astype(F))     # data/imagenet.py:16
        t = torch.from_numpy(rng.integers(0, 10, 32))
        lg = torch.nn.functional.cross_entropy(gpu(x.cuda()), t.cuda())
        lc = torch.nn.functional.cross_entropy(cpu(x), t)
        og.zero_grad(); oc.zero_grad()
        lg.backward(); lc.backward()
        assert abs(lg.item() - lc.item()) <= 2e-3 * max(1.0, abs(lc.item())), (step, lg.item(), lc.item())
        np.testing.assert_allclose(gpu.stage1_conv.weight.grad.cpu().numpy(), cpu.w1.grad.numpy(), rtol=0.05, atol=2e-3)
        og.step(); oc.step()
    _, aux, state = export_mx_params(gpu)
    suffix = ("_weight_quant", "_data_quant") if style == "quant_ops" else ("_weight", "_data")
    names = ["stage1_conv", "stage2_conv", "fc1"]
    for i, n in enumerate(names):
        for j, suf in enumerate(suffix):
            got = aux[n + suf + "_minmax"].cpu().numpy()
            want = cpu.q[2 * i + j].aux
            np.testing.assert_allclose(got, want, rtol=2e-3), (n, suf)
    assert set(aux) == {n + s + "_minmax" for n in names for s in suffix}
    assert all(v["delay_quant"] == 0 for v in state.values())


@pytest.mark.gpu
def test_resnet_int8_cifar_trains():
    """resnet_int8 topology (symbol/resnet_int8.py) at CIFAR size: every conv / FC through the quantization nodes,
    loss falls over a few SGD steps and the aux thresholds are populated."""
    import torch
    from b200quant.harness import ResNetInt8, quant_nodes
    torch.manual_seed(3)
    dev = torch.device("cuda", 0)
    model = ResNetInt8(units=(1, 1, 1), filter_list=(16, 16, 32, 64), num_classes=10, bottle_neck=False,
                       dataset_type="cifar10").to(dev)
    assert len(quant_nodes(model)) == 2 * (1 + 3 + 3 + 3 + 1)   # conv0, three units of conv1+conv2+sc, fc1
    opt = torch.optim.SGD(model.parameters(), lr=0.05, momentum=0.9)
    x = torch.randn(64, 3, 32, 32, device=dev)
    y = torch.randint(0, 10, (64,), device=dev)
    losses = []
    for _ in range(25):
        opt.zero_grad()
        loss = torch.nn.functional.cross_entropy(model(x), y)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert all(l == l for l in losses)
    assert min(losses[-5:]) < 0.8 * losses[0], losses
    for node in quant_nodes(model):
        for a in node.aux_list():
            assert float(a.abs().max()) > 0
