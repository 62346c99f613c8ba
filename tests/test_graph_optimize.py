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


# ---- graph-level rewriting (Concat / Pooling / adds), reference-semantics merge_bn, layer twins -------------------
class GraphNet(nn.Module):
    """pool0 feeds two convolutions AND the residual add; the branches are concatenated."""

    def __init__(self):
        super().__init__()
        self.conv0 = nn.Conv2d(3, 8, 3, padding=1, bias=False)
        self.bn0 = nn.BatchNorm2d(8)
        self.pool = nn.MaxPool2d(2)
        self.a = nn.Conv2d(8, 4, 1, bias=False)
        self.b = nn.Conv2d(8, 4, 3, padding=1, bias=False)
        self.c = nn.Conv2d(8, 8, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(8)
        self.up = nn.ConvTranspose2d(8, 4, 2, stride=2, bias=False)
        self.fc = nn.Linear(4, 5)

    def forward(self, x):
        p = self.pool(torch.relu(self.bn0(self.conv0(x))))
        cat = torch.cat([self.a(p), self.b(p)], 1)
        y = self.c(self.bn1(cat)) + p
        return self.fc(self.up(y).mean(dim=(2, 3)))


PACT = {"quantize_op_name": "PACT", "init_value": 6.0, "attrs": {"nbits": "4"}}
GDRQ_W = {"quantize_op_name": "GDRQ", "attrs": {"nbits": "4", "is_weight": "True"}}
ALL_OPS = ("Convolution", "FullyConnected", "Deconvolution", "Concat", "Pooling", "add_n", "elemwise_add")


def test_fx_rewriter_quantizes_concat_pooling_and_add_inputs_once_per_producer():
    """core/graph_optimize.py:216-217,261-272 on a torch model: data nodes on every input of Concat / Pooling /
    elemwise_add, named after their producer, shared by all consumers of that producer."""
    import torch.fx as fx
    from b200quant.graph_optimize import QuantNode, QuantizedWeightOp, attach_quantize_node
    gm = attach_quantize_node(GraphNet(), GDRQ_W, PACT, quantized_op=ALL_OPS)
    assert isinstance(gm, fx.GraphModule)
    assert gm.quantized_op_counts == {"Convolution": 4, "FullyConnected": 1, "Deconvolution": 1, "Concat": 1, "Pooling": 1,
                                      "add_n": 0, "elemwise_add": 1}
    mods = dict(gm.named_modules())
    nodes = {n.name: n for n in gm.graph.nodes}
    qnodes = [n for n in gm.graph.nodes if n.op == "call_module" and isinstance(mods[n.target], QuantNode)]
    producers = [n.args[0].name for n in qnodes]
    assert len(producers) == len(set(producers))                               # one node per producer
    assert all(mods[n.target].var_name == n.args[0].name for n in qnodes)       # named like the producer (:168-195)
    assert len(nodes["pool_quant"].users) == 3                                  # a, b and the residual add share it
    assert [a.name for a in nodes["cat"].args[0]] == ["a_quant", "b_quant"]
    assert [a.name for a in nodes["add"].args] == ["c_quant", "pool_quant"]
    assert nodes["pool"].args[0].name == "relu_quant"
    for name in ("conv0", "a", "b", "c", "up", "fc"):
        assert isinstance(mods[name], QuantizedWeightOp) and mods[name].weight_quant.var_name == name + "_weight"
        assert mods[name].weight_quant.node.op_type == "GDRQ_PY"
    # skip counts apply per operator kind, in graph order
    gm2 = attach_quantize_node(GraphNet(), GDRQ_W, PACT, quantized_op=ALL_OPS,
                               skip_quantize_counts={"Convolution": 1, "Pooling": 1})
    mods2 = dict(gm2.named_modules())
    assert isinstance(mods2["conv0"], nn.Conv2d) and isinstance(mods2["a"], QuantizedWeightOp)
    assert "relu_quant" not in {n.name for n in gm2.graph.nodes}


def test_merge_bn_follows_the_dataflow_and_the_reference_semantics():
    """graph_optimize.py:37-112: only a BatchNorm in inference mode fed by a convolution is replaced, by a per-channel
    scale and shift (the convolution's weights are untouched); the result equals the BatchNorm it replaces."""
    from b200quant.graph_optimize import ChannelAffine, merge_bn
    torch.manual_seed(3)
    m = GraphNet().eval()
    with torch.no_grad():
        for bn in (m.bn0, m.bn1):
            bn.running_mean.normal_()
            bn.running_var.uniform_(0.5, 1.5)
            bn.weight.uniform_(0.5, 1.5)
            bn.bias.normal_()
    x = torch.randn(2, 3, 8, 8)
    want = m(x)
    w0 = m.conv0.weight.detach().clone()
    g = merge_bn(m)
    mods = dict(g.named_modules())
    assert isinstance(mods["bn0"], ChannelAffine)          # conv0 -> bn0
    assert isinstance(mods["bn1"], nn.BatchNorm2d)         # fed by a concat, not by a convolution: left alone (:71)
    assert mods["bn0"].gamma.shape == (1, 8, 1, 1)
    assert torch.equal(mods["conv0"].weight, w0)
    torch.testing.assert_close(g(x), want, rtol=1e-5, atol=1e-5)
    # training-mode BatchNorm (batch statistics) is never folded; affine=False is handled
    t = nn.Sequential(nn.Conv2d(3, 4, 1), nn.BatchNorm2d(4)).train()
    assert isinstance(dict(merge_bn(t).named_modules())["1"], nn.BatchNorm2d)
    e = nn.Sequential(nn.Conv2d(3, 4, 1), nn.BatchNorm2d(4, affine=False)).eval()
    y = e(x)
    torch.testing.assert_close(merge_bn(e)(x), y, rtol=1e-5, atol=1e-5)
    with pytest.raises(AssertionError):
        merge_bn(nn.Sequential(nn.Conv2d(3, 4, 1), nn.BatchNorm2d(5)).eval())


def test_layer_twins_keep_the_int8_api_names():
    """symbol/int8_api.py:73-117: deconv / data / add / concat wrappers -- node names, aux names, weight layout."""
    from b200quant.harness import QuantAdd, QuantConcat, QuantData, QuantDeconv2d, export_mx_params
    net = nn.ModuleDict(dict(up=QuantDeconv2d("up0", 8, (2, 2), (2, 2), (0, 0), 6), d=QuantData("in0", delay_quant=2),
                             add=QuantAdd("res0"), cat=QuantConcat("cat0", 3)))
    assert tuple(net["up"].weight.shape) == (8, 6, 2, 2)          # (in_channels, num_filter, kh, kw) (:82-84)
    assert (net["up"].weight_node_name, net["up"].data_node_name) == ("up0_weight", "up0_data")
    assert net["d"].data_quant.node_name == "in0_data" and net["d"].data_quant.op.delay_quant == 2
    assert (net["add"].lhs_quant.node_name, net["add"].rhs_quant.node_name, net["add"].out_name) == \
        ("res0add_lhs_data", "res0add_rhs_data", "res0_plus")
    assert [q.node_name for q in net["cat"].quants] == ["cat0concat_0_data", "cat0concat_1_data", "cat0concat_2_data"]
    assert all(q.op_type == "ClipGrad_Quantization_int8" and not q.op.is_weight for q in net["cat"].quants)
    _, _, state = export_mx_params(net)
    assert state["in0_data"]["delay_quant"] == 2 and "res0add_lhs_data" in state


@pytest.mark.gpu
def test_fx_rewritten_model_and_layer_twins_train():
    from b200quant.graph_optimize import attach_quantize_node, export_quant_params
    from b200quant.harness import QuantAdd, QuantConcat, QuantData, QuantDeconv2d, export_mx_params
    torch.manual_seed(1)
    gm = attach_quantize_node(GraphNet(), GDRQ_W, PACT, quantized_op=ALL_OPS).cuda().train()
    x = torch.rand(4, 3, 8, 8, device="cuda") * 2 - 1
    loss = gm(x).square().mean()
    loss.backward()
    assert torch.isfinite(loss)
    args, aux = export_quant_params(gm)
    assert {"pool_gamma", "relu_gamma", "a_gamma", "b_gamma", "c_gamma", "x_gamma"} <= set(args)
    assert {"conv0_weight_alpha", "up_weight_alpha", "fc_weight_alpha"} <= set(aux)
    assert all(p.grad is not None for n, p in gm.named_parameters() if n.endswith("gamma") and "quant" in n)

    class Twins(nn.Module):
        def __init__(self):
            super().__init__()
            self.d = QuantData("in0")
            self.add = QuantAdd("res0")
            self.cat = QuantConcat("cat0", 2)
            self.up = QuantDeconv2d("up0", 6, (2, 2), (2, 2), (0, 0), 4)

        def forward(self, x):
            q = self.d(x)
            s = self.add(q, x * 0.5)
            return self.up(self.cat([s, q]))

    t = Twins().cuda().train()
    xin = (torch.rand(2, 3, 6, 6, device="cuda") * 4 - 2).requires_grad_(True)
    out = t(xin)
    assert out.shape == (2, 4, 12, 12)
    out.sum().backward()
    assert xin.grad is not None and torch.isfinite(xin.grad).all() and t.up.weight.grad is not None
    _, aux, _ = export_mx_params(t)
    assert {"in0_data_minmax", "res0add_lhs_data_minmax", "res0add_rhs_data_minmax", "cat0concat_0_data_minmax",
            "cat0concat_1_data_minmax", "up0_weight_minmax", "up0_data_minmax"} <= set(aux)
    # first batch: ClipGrad activations initialise the threshold with max|x| (clip_grad_quantization_int8.py:42-44)
    assert float(aux["in0_data_minmax"]) == float(xin.detach().abs().max())
