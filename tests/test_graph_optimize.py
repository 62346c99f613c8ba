"""attach_quantize_node (core/graph_optimize.py:199-292 equivalent): structure on CPU, execution on GPU."""
import pytest
import torch
import torch.nn as nn

SETTING = {
    "weight": {"quantize_op_name": "Quantization_int8", "init_value": 0,
               "attrs": {"nbits": "3", "quant_mode": "minmax", "is_weight": "True", "is_weight_perchannel": "False",
                         "delay_quant": "0", "ema_decay": "0.99", "grad_mode": "ste", "fix_act_scale": "False"}},
    "act": {"quantize_op_name": "Quantization_int8", "init_value": 0,
            "attrs": {"nbits": "4", "quant_mode": "minmax", "is_weight": "False", "is_weight_perchannel": "False",
                      "delay_quant": "0", "ema_decay": "0.99", "grad_mode": "ste", "fix_act_scale": "False"}},
}   # config/edict_config.py:158-187


class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv0 = nn.Conv2d(3, 8, 3, padding=1, bias=False)
        self.bn0 = nn.BatchNorm2d(8)
        self.branch_a = nn.Conv2d(8, 8, 1, bias=False)
        self.branch_b = nn.Conv2d(8, 8, 3, padding=1, bias=False)   # same input as branch_a
        self.fc0 = nn.Linear(8, 6)
        self.fc1 = nn.Linear(6, 4)

    def forward(self, x):
        x = torch.relu(self.bn0(self.conv0(x)))
        x = self.branch_a(x) + self.branch_b(x)
        x = x.mean(dim=(2, 3))
        return self.fc1(torch.relu(self.fc0(x)))


def test_structure_skip_counts_and_names():
    from b200quant.graph_optimize import QuantizedOp, QuantNode, attach_quantize_node
    m = attach_quantize_node(Net(), SETTING["weight"], SETTING["act"],
                             skip_quantize_counts={"Convolution": 1, "FullyConnected": 1})   # edict_config.py:191
    assert isinstance(m.conv0, nn.Conv2d) and isinstance(m.fc0, nn.Linear)          # first of each kind skipped
    for name in ("branch_a", "branch_b", "fc1"):
        q = getattr(m, name)
        assert isinstance(q, QuantizedOp)
        assert q.data_quant.node.op_type == "Quantization_int8_V2" and not q.data_quant.node.op.is_weight
        assert q.weight_quant.node.op.is_weight
        assert q.data_quant.var_name == name + "_data" and q.weight_quant.var_name == name + "_weight"
        assert q.data_quant.node.aux_names == ["minmax"]
    assert m.quantized_op_counts == {"Convolution": 3, "FullyConnected": 2, "Deconvolution": 0}
    assert sum(isinstance(x, QuantNode) for x in m.modules()) == 6


def test_create_quant_node_mapping():
    from b200quant.graph_optimize import create_quant_node
    pact = create_quant_node("relu1", {"quantize_op_name": "PACT", "init_value": 8.0, "attrs": {"nbits": "4"}})
    assert pact.node.op_type == "PACT_PY" and pact.param_names == ["gamma"] and float(pact.gamma) == 8.0
    qil = create_quant_node("w", {"quantize_op_name": "QIL", "attrs": {"is_weight": "True", "fix_gamma": "True", "nbits": "4"}})
    assert qil.param_names == ["pruning_point", "clipping_point", "gamma"]
    assert qil.pruning_point.lr_mult == 0.01 and qil.pruning_point.wd_mult == 0.0 and float(qil.clipping_point) == 1.0
    assert not qil.gamma.requires_grad
    gd = create_quant_node("x", {"quantize_op_name": "GDRQ", "init_value": 0.5,
                                 "attrs": {"nbits": "4", "fix_alpha": "False", "group_size": "-1", "is_weight": "True",
                                           "lamda": "0.001", "delay_quant": "0", "ktimes": "3"}})
    assert gd.node.op_type == "GDRQ_PY" and gd.node.aux_init == 0.5
    with pytest.warns(UserWarning, match="parity with the C\\+\\+ operator is unpinned"):
        cxx = create_quant_node("x", {"quantize_op_name": "GDRQ_CXX", "attrs": {"nbits": "8", "is_weight": "False"}})
    assert cxx.node.op_type == "GDRQ_PY"
    with pytest.warns(UserWarning):
        assert create_quant_node("r", {"quantize_op_name": "PACT_CXX", "attrs": {}}).param_names == ["gamma"]
    with pytest.raises(AssertionError):
        create_quant_node("x", {"quantize_op_name": "nope", "attrs": {}})


@pytest.mark.gpu
def test_rewritten_model_trains_and_dedups_shared_input():
    from b200quant.graph_optimize import attach_quantize_node, export_quant_params
    torch.manual_seed(0)
    m = attach_quantize_node(Net(), SETTING["weight"], SETTING["act"], skip_quantize_counts={"Convolution": 1}).cuda().train()
    x = torch.rand(4, 3, 8, 8, device="cuda") * 2 - 1
    loss = m(x).square().mean()
    loss.backward()
    assert torch.isfinite(loss)
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    _, aux = export_quant_params(m)
    # branch_a and branch_b read the same tensor: quantized once, by the first consumer (graph_optimize.py:247-258)
    assert "branch_a_data_minmax" in aux and "branch_b_data_minmax" not in aux
    assert float(aux["branch_a_data_minmax"]) != 0.0
    assert {"branch_a_weight_minmax", "branch_b_weight_minmax", "fc0_data_minmax", "fc1_weight_minmax"} <= set(aux)


@pytest.mark.gpu
def test_pact_and_qil_nodes_learn_their_scalars():
    from b200quant.graph_optimize import attach_quantize_node
    act = {"quantize_op_name": "PACT", "init_value": 1.5, "attrs": {"nbits": "4"}}
    wgt = {"quantize_op_name": "QIL", "init_value": 0.9, "attrs": {"is_weight": "True", "fix_gamma": "True", "nbits": "4"}}
    net = nn.Sequential(nn.Conv2d(3, 4, 3, padding=1, bias=False), nn.ReLU(), nn.Conv2d(4, 4, 3, padding=1, bias=False))
    with torch.no_grad():
        for p in net.parameters():
            p.uniform_(-1, 1)
    m = attach_quantize_node(net, wgt, act).cuda().train()
    out = m(torch.rand(2, 3, 6, 6, device="cuda") * 3)
    out.sum().backward()
    q = m[2]
    assert q.data_quant.gamma.grad is not None and q.weight_quant.clipping_point.grad is not None
    assert q.weight_quant.gamma.grad is None


def test_quant_attrs_schema_validates_reference_settings():
    from b200quant.quant_attrs import validate, validate_quantize_setting
    validate_quantize_setting(SETTING)
    with pytest.raises(ValueError):
        validate({"quantize_op_name": "GDRQ", "attrs": {"nbits": 4}})            # not a string
    with pytest.raises(ValueError):
        validate({"quantize_op_name": "PACT", "attrs": {"group_size": "2"}})      # foreign attribute
    with pytest.raises(ValueError):
        validate({"quantize_op_name": "nope", "attrs": {}})
