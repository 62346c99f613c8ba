"""MXNet .params container: round trip, byte layout of a known tiny file, and restoring a quantized torch model."""
import struct

import numpy as np
import pytest


def test_round_trip_and_layout(tmp_path):
    from b200quant import checkpoint as ck
    rng = np.random.default_rng(0)
    named = {"arg:conv0_weight": rng.standard_normal((4, 3, 3, 3)).astype(np.float32),
             "aux:conv0_data_quant_minmax": np.array([1.25], np.float32),
             "aux:bn0_moving_var": rng.uniform(0.5, 1.5, 4).astype(np.float32),
             "arg:labels": np.arange(5, dtype=np.int64)}
    p = str(tmp_path / "x.params")
    ck.save_ndarray_dict(p, named)
    back = ck.load_ndarray_dict(p)
    assert list(back) == list(named)
    for k in named:
        assert back[k].dtype == named[k].dtype and np.array_equal(back[k], named[k])
    raw = open(p, "rb").read()
    assert struct.unpack("<QQQ", raw[:24]) == (0x112, 0, 4)                       # list magic, reserved, count
    assert struct.unpack("<IiI", raw[24:36]) == (0xF993FAC9, 0, 4)                # V2 magic, dense stype, ndim
    assert struct.unpack("<4q", raw[36:68]) == (4, 3, 3, 3)                       # int64 dims
    assert struct.unpack("<iii", raw[68:80]) == (1, 0, 0)                         # cpu(0), float32


def test_reads_v1_entries(tmp_path):
    """older files: V1 magic, uint32 dims, no storage type"""
    from b200quant import checkpoint as ck
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    p = str(tmp_path / "v1.params")
    with open(p, "wb") as f:
        f.write(struct.pack("<QQQ", 0x112, 0, 1))
        f.write(struct.pack("<II2I", 0xF993FAC8, 2, 2, 3))
        f.write(struct.pack("<iii", 1, 0, 0))
        f.write(a.tobytes())
        f.write(struct.pack("<QQ", 1, 5) + b"arg:w")
    assert np.array_equal(ck.load_ndarray_dict(p)["arg:w"], a)


def test_checkpoint_names_and_restore(tmp_path):
    import torch
    from b200quant import checkpoint as ck
    from b200quant.harness import SimpleCifarNet, export_mx_params
    torch.manual_seed(1)
    m = SimpleCifarNet()
    # give the quant nodes their aux buffers without running kernels (no GPU here)
    for mod in m.modules():
        if hasattr(mod, "aux_names") and mod.aux_names:
            mod.register_buffer("minmax", torch.full((1,), 0.5 + torch.rand(1).item()))
            mod._aux_ready = True
    m.stage1_conv.data_quant.op.delay_quant = 3
    m.stage1_conv.data_quant.op.init = False
    args, aux, state = export_mx_params(m)
    assert "stage1_conv_data_quant_minmax" in aux and "fc1_weight_quant_minmax" in aux
    path = ck.save_checkpoint(str(tmp_path / "net"), 7, args, aux, state)
    assert path.endswith("net-0007.params")
    a2, x2, s2 = ck.load_checkpoint(str(tmp_path / "net"), 7)
    assert set(a2) == set(args) and set(x2) == set(aux)
    m2 = SimpleCifarNet()
    missing = ck.restore_quant_layers(m2, a2, x2, s2)
    assert missing == []
    assert torch.equal(m2.stage2_conv.weight, m.stage2_conv.weight)
    assert torch.equal(m2.fc1.data_quant.minmax, m.fc1.data_quant.minmax)
    assert m2.stage1_conv.data_quant.op.delay_quant == 3 and m2.stage1_conv.data_quant.op.init is False
    # an fp32 checkpoint (no thresholds) seeds a quantized graph: allow_missing like the reference (edict_config.py:27)
    m3 = SimpleCifarNet()
    missing = ck.restore_quant_layers(m3, a2, {}, None)
    assert "stage1_conv_data_quant_minmax" in missing and torch.equal(m3.fc1.weight, m.fc1.weight)
    with pytest.raises(KeyError):
        ck.restore_quant_layers(SimpleCifarNet(), a2, {}, None, allow_missing=False)
