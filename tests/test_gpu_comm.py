"""One process, N devices (include/b2q.h b2q_comm_*; the reference's layout, train.py:34 / solver.py:58-61): the fused
threshold exchange and the slice-owner allreduce over peer memory, without torch.distributed.  Needs >= 2 GPUs."""
import numpy as np
import pytest

from oracle import quant_oracle as qo
from tests.golden_util import bits_equal

pytestmark = pytest.mark.gpu
F = np.float32


@pytest.fixture(scope="module")
def group():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    from b200quant.comm import DeviceGroup
    n = min(torch.cuda.device_count(), 4)
    g = DeviceGroup(list(range(n)))
    yield g
    for d in range(n):
        torch.cuda.synchronize(d)
    g.close()


@pytest.mark.parametrize("count", [1, 7, 4096, 1000003, 25_502_912])
def test_allreduce_sum_and_max_are_exact_and_identical_on_every_rank(group, count):
    import torch
    xs = [torch.empty(count, device="cuda:%d" % d).normal_(generator=torch.Generator(device="cuda:%d" % d).manual_seed(d))
          for d in group.devices]
    host = [x.cpu().numpy() for x in xs]
    want = host[0].copy()
    for h in host[1:]:
        want = (want + h).astype(F)                       # rank order, float32 adds: what the slice owner computes
    group.allreduce_sum(xs)
    for d in group.devices:
        torch.cuda.synchronize(d)
    for x in xs:
        assert bits_equal(x.cpu().numpy(), want)
    ys = [torch.from_numpy(h).to("cuda:%d" % d) for h, d in zip(host, group.devices)]
    group.allreduce_sum(ys, average=True)
    zs = [torch.from_numpy(h).to("cuda:%d" % d) for h, d in zip(host, group.devices)]
    group.allreduce_max(zs)
    for d in group.devices:
        torch.cuda.synchronize(d)
    want_avg = (want * F(1.0 / len(host))).astype(F)
    want_max = np.maximum.reduce(host)
    for y, z in zip(ys, zs):
        assert bits_equal(y.cpu().numpy(), want_avg) and bits_equal(z.cpu().numpy(), want_max)


@pytest.mark.parametrize("op_type", ["Quantization_int8_V2", "ClipGrad_Quantization_int8", "GDRQ_PY"])
def test_fused_threshold_exchange_in_one_process(group, op_type):
    """every rank quantises its own shard; aux is identical on all ranks and equals the oracle fed the max over ranks of the
    statistic; every rank's output is bit-identical to the oracle given that threshold."""
    import torch
    import b200quant
    if op_type == "GDRQ_PY":
        attrs = dict(nbits="8", group_size="-1", is_weight="False", lamda="0.001", delay_quant="0", fix_alpha="False", ktimes="3")
    else:
        attrs = dict(quant_mode="minmax", is_weight="False", is_weight_perchannel="False", delay_quant="0", ema_decay="0.99")
    ops = [b200quant.get_prop(op_type)(**attrs).create_operator(None, None, None) for _ in group.devices]
    group.attach_threshold_exchange([[op] for op in ops])
    refs = [qo.create(op_type, **attrs) for _ in group.devices]
    aux = [torch.ones(1, device="cuda:%d" % d) for d in group.devices]
    aux_ref = np.ones(1, F)
    for step in range(3):
        shape = (4, 16, 28, 28) if step != 1 else (3, 7, 5)
        xs = [(np.random.default_rng(100 * step + d).standard_normal(shape) * (1 + d)).astype(F) for d in group.devices]
        xd = [torch.from_numpy(x).to("cuda:%d" % d) for x, d in zip(xs, group.devices)]
        yd = [torch.zeros_like(x) for x in xd]
        for r, d in enumerate(group.devices):       # one host thread issues every rank's work, as MXNet's executor group does
            with torch.cuda.device(d):
                ops[r].forward(True, ["write"], [xd[r]], [yd[r]], [aux[r]])
        for d in group.devices:
            torch.cuda.synchronize(d)
        group.check()
        if op_type == "GDRQ_PY":
            stat = max(F(F(np.abs(x).sum(dtype=np.float64)) / F(x.size)) for x in xs)
            thr = F(F(3.0) * stat)
            aux_ref = np.array([F(aux_ref[0] + F(F(0.001) * F(aux_ref[0] - thr)))], F)
        else:
            stat = max(F(np.abs(x).max()) for x in xs)
            if op_type == "ClipGrad_Quantization_int8" and step == 0:
                aux_ref = np.array([stat], F)
            else:
                aux_ref = qo.mx_add(qo.mx_mul(aux_ref, F(0.99)), qo.mx_mul(stat, F(1 - 0.99))).reshape(1)
        for r in range(group.world):
            got = aux[r].cpu().numpy()
            if op_type == "GDRQ_PY":
                np.testing.assert_allclose(got, aux_ref, rtol=1e-6)
                assert bits_equal(got, aux[0].cpu().numpy())
                a = got.copy()
            else:
                assert bits_equal(got, aux_ref), (r, step)
                a = aux_ref.copy()
            # oracle sweep with the agreed threshold
            if op_type == "GDRQ_PY":
                c = qo.mx_clip(xs[r], -a[0], a[0])
                want, _ = qo.qdq(c, qo.mx_div(a, F(255)))
            elif op_type == "ClipGrad_Quantization_int8":
                want, _ = qo.qdq(qo.mx_clip(xs[r], -a[0], a[0]), qo.mx_div(a, F(127)))
            else:
                want, _ = qo.qdq(xs[r], qo.mx_div(a, F(127)))
            assert bits_equal(yd[r].cpu().numpy(), want), (r, step)
        if op_type == "GDRQ_PY":
            aux_ref = aux[0].cpu().numpy().copy()
