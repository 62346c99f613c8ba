"""Single-launch forward of small whole-tensor nodes on one thread-block cluster (csrc/b2q_cluster.cuh: the tensor staged
in the CTAs' shared memory, partial statistics exchanged through distributed shared memory between two cluster
barriers) against the two-kernel path and the oracle: every operator that routes through it, sizes around the 1/2/4/8
CTA steps and the eligibility limit, misaligned views, several training steps (EMA / first-batch state)."""
import numpy as np
import pytest

from oracle import quant_oracle as qo
from tests.golden_util import bits_equal

pytestmark = pytest.mark.gpu
F = np.float32
SIZES = [1, 7, 8, 9, 63, 4096, 8191, 8192, 8200, 9408, 16384, 36864, 65536, 65544, 147456, 262144, 327672, 327680]


def _ops(b200quant):
    mk = lambda t, **kw: (lambda: b200quant.get_prop(t)(**{k: str(v) for k, v in kw.items()}).create_operator(None, None, None))
    return {
        "v2_weight": (mk("Quantization_int8_V2", quant_mode="minmax", is_weight=True), 1),
        "v2_act": (mk("Quantization_int8_V2", quant_mode="minmax", is_weight=False), 1),
        "clipgrad_weight": (mk("ClipGrad_Quantization_int8", quant_mode="minmax", is_weight=True), 1),
        "clipgrad_act": (mk("ClipGrad_Quantization_int8", quant_mode="minmax", is_weight=False), 1),
        "gdrq_weight": (mk("GDRQ_PY", nbits=8, group_size=-1, is_weight=True), 1),
        "gdrq_act": (mk("GDRQ_PY", nbits=4, group_size=-1, is_weight=False), 1),
    }


@pytest.mark.parametrize("kind", ["v2_weight", "v2_act", "clipgrad_weight", "clipgrad_act", "gdrq_weight", "gdrq_act"])
def test_cluster_forward_equals_two_kernel_path(kind):
    import torch
    import b200quant
    from b200quant import _lib
    ctx = _lib.context(0)
    saved_opts = {k: ctx.get_option(k) for k in ("cluster_max_elems", "cluster_max_elems_mean")}
    saved = ctx.get_option("cluster_fwd")
    g = torch.Generator(device="cuda").manual_seed(13)
    mk, naux = _ops(b200quant)[kind]
    try:
        ctx.set_option("cluster_max_elems", 327680)          # the kernel's own limit (the defaults are the measured
        ctx.set_option("cluster_max_elems_mean", 327680)     # crossovers against two launches, well below it)
        a, b = mk(), mk()
        aux_a, aux_b = torch.ones(naux, device="cuda"), torch.ones(naux, device="cuda")
        for step, n in enumerate(SIZES):
            for off in (0, 3):
                buf = torch.randn(n + 8, device="cuda", generator=g) * (0.05 + 0.3 * (step % 4))
                x = buf[off:off + n]
                ya, yb = torch.full((n + 8,), 7.0, device="cuda"), torch.full((n + 8,), 7.0, device="cuda")
                ctx.set_option("cluster_fwd", 0)
                l0 = ctx.launch_count()
                a.forward(True, ["write"], [x], [ya[off:off + n]], [aux_a])
                two = ctx.launch_count() - l0
                ctx.set_option("cluster_fwd", 1)
                l0 = ctx.launch_count()
                b.forward(True, ["write"], [x], [yb[off:off + n]], [aux_b])
                one = ctx.launch_count() - l0
                if n >= 63:
                    assert one == 1 and two == 2, (kind, n, off, one, two)
                if kind.startswith("gdrq"):     # mean-based: a differently partitioned (exact, double) sum
                    torch.testing.assert_close(aux_a, aux_b, rtol=2e-7, atol=0)
                    aux_a.copy_(aux_b)
                else:
                    assert torch.equal(aux_a.view(torch.int32), aux_b.view(torch.int32)), (kind, n, off)
                    assert torch.equal(ya.view(torch.int32), yb.view(torch.int32)), (kind, n, off)
                assert bool((yb[:off] == 7.0).all()) and bool((yb[off + n:] == 7.0).all())
        # above the limit the cluster path steps aside
        x = torch.randn(327681 + 8, device="cuda", generator=g)[:327681]
        y = torch.empty_like(x)
        l0 = ctx.launch_count()
        b.forward(True, ["write"], [x], [y], [aux_b])
        assert ctx.launch_count() - l0 == 2
    finally:
        ctx.set_option("cluster_fwd", saved)
        for k, v in saved_opts.items():
            ctx.set_option(k, v)


@pytest.mark.parametrize("op_type,attrs", [
    ("Quantization_int8_V2", dict(quant_mode="minmax", is_weight="True")),
    ("Quantization_int8_V2", dict(quant_mode="minmax", is_weight="False")),
    ("ClipGrad_Quantization_int8", dict(quant_mode="minmax", is_weight="False")),
    ("ClipGrad_Quantization_int8", dict(quant_mode="minmax", is_weight="True")),
    ("GDRQ_PY", dict(nbits="8", group_size="-1", is_weight="True")),
    ("GDRQ_PY", dict(nbits="8", group_size="-1", is_weight="False", lamda="0.001", ktimes="3")),
])
def test_cluster_forward_equals_oracle_over_training_steps(op_type, attrs):
    """three training steps and one evaluation step per size through the CustomOp protocol, compared bit for bit with
    the NumPy oracle (quant_ops.py:12-40, clip_grad_quantization_int8.py:14-53, GDRQ.py:64-86)."""
    import torch
    import b200quant
    from b200quant import _lib
    ctx = _lib.context(0)
    assert ctx.get_option("cluster_fwd") == 1
    saved = ctx.get_option("cluster_max_elems")
    ctx.set_option("cluster_max_elems", 65536)       # the max-based operators too (off by default: neutral in a step)
    try:
        _oracle_steps(ctx, op_type, attrs)
    finally:
        ctx.set_option("cluster_max_elems", saved)


def _oracle_steps(ctx, op_type, attrs):
    import torch
    import b200quant
    rng = np.random.default_rng(23)
    for shape in [(64, 64, 1, 1), (64, 3, 7, 7), (128, 128, 3, 3), (1024, 256, 1, 1), (32, 3, 32, 32), (5, 7, 3)]:
        op = b200quant.get_prop(op_type)(**attrs).create_operator(None, None, None)
        ref = qo.create(op_type, **attrs)
        aux, aux_r = torch.ones(1, device="cuda"), np.ones(1, F)
        for step in range(4):
            is_train = step < 3
            x = (rng.standard_normal(shape) * (0.2 + step)).astype(F)
            xd = torch.from_numpy(x).cuda()
            y = torch.zeros_like(xd)
            l0 = ctx.launch_count()
            op.forward(is_train, ["write"], [xd], [y], [aux])
            launches = ctx.launch_count() - l0
            yr = np.zeros_like(x)
            ref.forward(is_train, ["write"], [x], [yr], [aux_r])
            assert bits_equal(aux.cpu().numpy(), aux_r), (shape, step)
            assert bits_equal(y.cpu().numpy(), yr), (shape, step)
            reduces = is_train or op_type == "GDRQ_PY" or (op_type == "Quantization_int8_V2" and attrs["is_weight"] == "True")
            n = int(np.prod(shape))
            limit = 147456 if op_type == "GDRQ_PY" else 65536
            assert launches == (1 if (n <= limit or not reduces) else 2), (shape, step, launches, reduces)


def test_cluster_forward_special_values():
    """NaN / Inf / all-zero / denormal tensors behave as on the two-kernel path (NaN propagates, q = 0 gives NaN)."""
    import torch
    import b200quant
    from b200quant import _lib
    ctx = _lib.context(0)
    saved = ctx.get_option("cluster_fwd")
    saved_max = ctx.get_option("cluster_max_elems")
    ctx.set_option("cluster_max_elems", 65536)
    cases = {"zeros": np.zeros(5000, F), "nan": np.r_[np.ones(4999, F), F(np.nan)], "inf": np.r_[np.ones(4999, F), F(np.inf)],
             "denormal": np.full(5000, 1e-40, F), "outlier": np.r_[np.full(4999, 1e-3, F), F(1e30)]}
    try:
        for name, x in cases.items():
            for op_type, is_w in (("Quantization_int8_V2", "True"), ("ClipGrad_Quantization_int8", "False"), ("GDRQ_PY", "False")):
                outs = []
                for mode in (0, 1):
                    ctx.set_option("cluster_fwd", mode)
                    kw = dict(quant_mode="minmax", is_weight=is_w) if op_type != "GDRQ_PY" else dict(nbits="8", group_size="-1", is_weight=is_w)
                    op = b200quant.get_prop(op_type)(**kw).create_operator(None, None, None)
                    xd = torch.from_numpy(x.astype(F)).cuda()
                    y, aux = torch.zeros_like(xd), torch.ones(1, device="cuda")
                    op.forward(True, ["write"], [xd], [y], [aux])
                    outs.append((y.cpu().numpy(), aux.cpu().numpy()))
                assert bits_equal(outs[0][0], outs[1][0]) and bits_equal(outs[0][1], outs[1][1]), (name, op_type)
    finally:
        ctx.set_option("cluster_fwd", saved)
        ctx.set_option("cluster_max_elems", saved_max)
