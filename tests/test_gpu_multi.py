"""Multi-tensor weight launches must be bit-identical to driving the weight operators one by one."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
F = np.float32

SHAPES = [((64, 3, 7, 7), False), ((64, 64, 1, 1), False), ((512, 512, 3, 3), False), ((1000, 2048), False),
          ((32, 1, 3, 3), True), ((1024, 1, 3, 3), True), ((128, 64, 1, 1), True), ((8, 4100), True),
          ((10, 8), True), ((5, 3, 1, 1), False), ((7, 11, 3, 3), True), ((3, 1030), False)]


@pytest.mark.parametrize("op_type", ["Quantization_int8_V2", "ClipGrad_Quantization_int8"])
def test_weight_group_equals_per_op(op_type):
    import torch
    import b200quant
    from b200quant.multi import WeightGroup

    def mk_ops():
        return [b200quant.get_prop(op_type)(quant_mode="minmax", is_weight="True", is_weight_perchannel=str(pc),
                                            delay_quant="0").create_operator(None, None, None) for _, pc in SHAPES]

    g = torch.Generator(device="cuda").manual_seed(5)
    ops_a, ops_b = mk_ops(), mk_ops()
    ws = [torch.randn(s, device="cuda", generator=g) * (0.05 + 0.1 * i) for i, (s, _) in enumerate(SHAPES)]
    dys = [torch.randn(s, device="cuda", generator=g) for s, _ in SHAPES]
    ya = [torch.zeros_like(w) for w in ws]
    yb = [torch.zeros_like(w) for w in ws]
    dxa = [torch.zeros_like(w) for w in ws]
    dxb = [torch.zeros_like(w) for w in ws]
    auxa = [torch.ones(s[0] if pc else 1, device="cuda") for s, pc in SHAPES]
    auxb = [a.clone() for a in auxa]
    group = WeightGroup(ops_b, ws, yb, auxb, dys, dxb)
    for step, train in enumerate([True, False, True, True]):
        for w in ws:
            w.mul_(1.0 + 0.3 * step)
        for op, w, y, a in zip(ops_a, ws, ya, auxa):
            op.forward(train, ["write"], [w], [y], [a])
        group.forward(train)
        for op, w, y, a, dy, dx in zip(ops_a, ws, ya, auxa, dys, dxa):
            op.backward(["write"], [dy], [w], [y], [dx], [a])
        group.backward()
        torch.cuda.synchronize()
        for i in range(len(ws)):
            assert torch.equal(auxa[i].view(torch.int32), auxb[i].view(torch.int32)), (step, i, "aux")
            assert torch.equal(ya[i].view(torch.int32), yb[i].view(torch.int32)), (step, i, "y")
            assert torch.equal(dxa[i].view(torch.int32), dxb[i].view(torch.int32)), (step, i, "dx")
    group.close()


def test_weight_group_respects_delay_quant():
    import torch
    import b200quant
    from b200quant.multi import WeightGroup
    ops = [b200quant.get_prop("Quantization_int8_V2")(quant_mode="minmax", is_weight="True", delay_quant="1")
           .create_operator(None, None, None) for _ in range(2)]
    ws = [torch.randn(8, 4, 3, 3, device="cuda") for _ in range(2)]
    ys = [torch.zeros_like(w) for w in ws]
    auxs = [torch.ones(1, device="cuda") for _ in range(2)]
    group = WeightGroup(ops, ws, ys, auxs)
    group.forward(True)                       # delay step: pass-through, aux untouched
    assert all(torch.equal(y, w) for y, w in zip(ys, ws)) and all(float(a) == 1.0 for a in auxs)
    group.forward(True)                       # quantised now
    assert all(float(a) == float(w.abs().max()) for a, w in zip(auxs, ws))
    assert not torch.equal(ys[0], ws[0])


@pytest.mark.parametrize("fix_alpha", [False, True])
def test_gdrq_weight_group_equals_per_op(fix_alpha):
    import torch
    import b200quant
    from b200quant.multi import WeightGroup
    shapes = [((64, 3, 7, 7), -1), ((256, 64, 1, 1), -1), ((32, 1, 3, 3), 4), ((512, 512, 3, 3), -1), ((64, 32, 3, 3), 8),
              ((16, 1200), 2), ((1000, 1024), -1)]

    def mk():
        return [b200quant.get_prop("GDRQ_PY")(nbits="8", group_size=str(gs), is_weight="True", lamda="0.001",
                                               delay_quant="1", fix_alpha=str(fix_alpha), ktimes="3")
                .create_operator(None, None, None) for _, gs in shapes]

    g = torch.Generator(device="cuda").manual_seed(7)
    ops_a, ops_b = mk(), mk()
    ws = [torch.randn(s, device="cuda", generator=g) * 0.1 for s, _ in shapes]
    ya, yb = [torch.zeros_like(w) for w in ws], [torch.zeros_like(w) for w in ws]
    auxa = [torch.full((1 if gs == -1 else s[0] // gs,), 0.2, device="cuda") for s, gs in shapes]
    auxb = [a.clone() for a in auxa]
    group = WeightGroup(ops_b, ws, yb, auxb)
    for step in range(3):                       # step 0 runs the delay_quant (clip only) branch
        for w in ws:
            w.mul_(1.0 + 0.5 * step)
        for op, w, y, a in zip(ops_a, ws, ya, auxa):
            op.forward(True, ["write"], [w], [y], [a])
        group.forward(True)
        torch.cuda.synchronize()
        for i in range(len(ws)):
            assert torch.allclose(auxa[i], auxb[i], rtol=1e-6, atol=0), (step, i)
            if torch.equal(auxa[i].view(torch.int32), auxb[i].view(torch.int32)):
                assert torch.equal(ya[i].view(torch.int32), yb[i].view(torch.int32)), (step, i)
            else:  # summation order differs between the two schedules: thresholds may differ in the last ulp
                assert torch.allclose(ya[i], yb[i], rtol=1e-5, atol=1e-6), (step, i)
    group.close()
