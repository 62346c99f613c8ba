"""N>1 host logic on CPU: world_size-2 gloo.  The kernels cannot run here, so the per-rank statistic comes from
the oracle; what is checked is the exchange itself: after ThresholdSync every rank holds max over ranks, the
threshold every rank then derives equals the oracle fed that max, and GradBucket sums gradients in one call."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

F = np.float32


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import b200quant  # noqa: F401
        from b200quant.dist import GradBucket, ThresholdSync, attach_threshold_sync
        from oracle import quant_oracle as qo
        rng = np.random.default_rng(5 + rank)          # per-rank shard, seeds 5 + rank
        x = (rng.standard_normal((4, 8, 6, 6)) * (1.0 + rank)).astype(F)
        local = float(qo.mx_absmax(x))
        stat = torch.tensor([local])
        sync = ThresholdSync()
        sync(stat)
        # every rank derives the same EMA threshold from the synchronised statistic
        aux = qo.mx_add(qo.mx_mul(F(1.0), F(0.99)), qo.mx_mul(F(stat.item()), F(1 - 0.99)))
        # operators pick the sync up only where it belongs (activation nodes)
        a = b200quant.get_prop("Quantization_int8_V2")(quant_mode="minmax", is_weight="False").create_operator(None, None, None)
        w = b200quant.get_prop("Quantization_int8_V2")(quant_mode="minmax", is_weight="True").create_operator(None, None, None)
        attach_threshold_sync([a, w])
        bucket = GradBucket([(2, 3), (4,)], torch.device("cpu"))
        bucket.views[0].fill_(float(rank + 1))
        bucket.views[1].fill_(10.0 * (rank + 1))
        bucket.allreduce(average=False)
        q.put((rank, local, float(stat.item()), float(aux), a.sync is not None, w.sync is None,
               bucket.flat.tolist(), sync.calls))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_threshold_sync_and_grad_bucket_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    locals_ = [r[1] for r in res]
    assert locals_[0] != locals_[1]
    for r in res:
        assert r[2] == max(locals_)                       # allreduce(max)
        assert r[3] == res[0][3]                          # identical threshold on every rank
        assert r[4] and r[5]                              # sync attached to activation node only
        assert r[6] == [3.0] * 6 + [30.0] * 4             # allreduce(sum) over the flat bucket
        assert r[7] == 1


def test_ranks_of_a_node_get_disjoint_host_cores():
    """dist.pin_rank_to_host_cores (the e2e leg at N > 1): without a NUMA hint every local rank gets its own slice of the
    cores the process may use, the slices are disjoint, and the copy-pool size follows.  Run in child processes: the
    call changes the caller's CPU affinity."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ncpu = len(os.sched_getaffinity(0))
    world = 2 if ncpu >= 2 else 1
    got = []
    for r in range(world):
        code = ("import sys, os, json; sys.path.insert(0, %r); import b200quant; "
                "from b200quant.dist import pin_rank_to_host_cores; "
                "c = pin_rank_to_host_cores(%d, %d); "
                "print(json.dumps([c, sorted(os.sched_getaffinity(0)), os.environ.get('B2Q_HOST_COPY_THREADS')]))"
                % (root, r, world))
        env = dict(os.environ)
        env.pop("B2Q_HOST_COPY_THREADS", None)
        out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
        assert out.returncode == 0, out.stderr[-2000:]
        got.append(json.loads(out.stdout.strip().splitlines()[-1]))
    for cores, affinity, threads in got:
        assert cores == affinity and len(cores) >= 1 and int(threads) == len(cores)
    if world == 2:
        assert not set(got[0][0]) & set(got[1][0])
        assert len(got[0][0]) == ncpu // 2
