"""Multi-GPU parity check, run under torchrun (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py

Every rank quantises its own shard; thresholds are exchanged (a) with NCCL allreduce(max) and (b) with the fused
peer-memory kernels.  Rank r checks, bit for bit, against the NumPy oracle fed the max over ranks: aux after every
step and its own quantised output.  Exit code 0 = all good."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
F = np.float32


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import b200quant
    from b200quant.dist import ThresholdSync, attach_peer_exchange
    from oracle import quant_oracle as qo

    def bits(a, b):
        return np.array_equal(np.ascontiguousarray(a, dtype=F).view(np.uint32), np.ascontiguousarray(b, dtype=F).view(np.uint32))

    failures = 0
    from b200quant import _lib
    ctx = _lib.context(local)
    for k, v in os.environ.items():   # B2Q_OPT_<option>=<int>, e.g. B2Q_OPT_PEER_MODE=2 (results never depend on them)
        if k.startswith("B2Q_OPT_"):
            ctx.set_option(k[len("B2Q_OPT_"):].lower(), int(v))
    for mode in ("nccl", "peer"):
        for op_type, variant in (("Quantization_int8_V2", 0), ("ClipGrad_Quantization_int8", 1)):
            op = b200quant.get_prop(op_type)(quant_mode="minmax", is_weight="False").create_operator(None, None, None)
            ex = None
            if mode == "nccl":
                op.sync = ThresholdSync()
            else:
                ex = attach_peer_exchange([op], torch.device("cuda", local))
            aux = torch.ones(1, device="cuda")
            aux_ref = np.ones(1, F)
            init = True
            for step in range(4):
                rng = np.random.default_rng(1000 * step + 5 + rank)
                shape = (4, 16, 28, 28) if step % 2 == 0 else (3, 7, 5)
                x = (rng.standard_normal(shape) * (1.0 + rank + 0.5 * step)).astype(F)
                xd = torch.from_numpy(x).cuda()
                yd = torch.zeros_like(xd)
                op.forward(True, ["write"], [xd], [yd], [aux])
                # oracle: statistic = max over ranks of the per-rank absmax
                m = torch.tensor([float(np.abs(x).max())], device="cuda")
                dist.all_reduce(m, op=dist.ReduceOp.MAX)
                stat = F(m.item())
                if variant == 1 and init:
                    aux_ref[...] = stat
                else:
                    aux_ref[...] = qo.mx_add(qo.mx_mul(aux_ref, F(0.99)), qo.mx_mul(stat, F(1 - 0.99)))
                init = False
                q = qo.mx_div(aux_ref, F(127))
                src = qo.mx_clip(x, -float(aux_ref[0]), float(aux_ref[0])) if variant == 1 else x
                want, _ = qo.qdq(src, q)
                ok = bits(aux.cpu().numpy(), aux_ref) and bits(yd.cpu().numpy(), want)
                if not ok:
                    failures += 1
                    print("rank %d FAIL mode=%s op=%s step=%d aux=%r want=%r" % (rank, mode, op_type, step,
                                                                                 aux.cpu().numpy(), aux_ref), flush=True)
            if ex is not None:
                torch.cuda.synchronize()
                dist.barrier()
                ex.close()
    # GDRQ_PY activations (mean-based threshold): statistic = max over ranks of mean|x|, then the alpha update
    for group_size, mode in ((-1, "nccl"), (4, "nccl"), (-1, "peer"), (4, "peer"), (1, "peer")):
        op = b200quant.get_prop("GDRQ_PY")(nbits="8", group_size=str(group_size), is_weight="False", lamda="0.001",
                                           ktimes="3").create_operator(None, None, None)
        ex = None
        if mode == "nccl":
            op.sync = ThresholdSync()
        else:
            ex = attach_peer_exchange([op], torch.device("cuda", local))
        groups = 1 if group_size == -1 else 16 // group_size
        alpha = torch.ones(groups, device="cuda")
        alpha_ref = np.ones(groups, F)
        for step in range(3):
            rng = np.random.default_rng(77 * step + 5 + rank)
            x = (rng.standard_normal((4, 16, 14, 14)) * (1.0 + rank + 0.5 * step)).astype(F)
            xd = torch.from_numpy(x).cuda()
            yd = torch.zeros_like(xd)
            op.forward(True, ["write"], [xd], [yd], [alpha])
            if group_size == -1:
                mean = np.atleast_1d(qo.mx_mean(np.abs(x))).astype(F)
                a_full = None
            else:   # channel-major groups of `group_size` channels (GDRQ.py:88-98)
                xg = np.ascontiguousarray(np.swapaxes(x, 0, 1)).reshape(groups, -1)
                mean = np.array([qo.mx_mean(np.abs(xg[g])) for g in range(groups)], F)
            m = torch.from_numpy(mean).cuda()
            dist.all_reduce(m, op=dist.ReduceOp.MAX)
            thr = qo.mx_mul(m.cpu().numpy(), F(3.0))
            alpha_ref = qo.mx_add(alpha_ref, qo.mx_mul(F(0.001), qo.mx_sub(alpha_ref, thr)))
            a_el = alpha_ref[0] if group_size == -1 else np.repeat(alpha_ref, group_size)[None, :, None, None]
            if group_size == -1:
                c = qo.mx_clip(x, -float(a_el), float(a_el))
            else:
                c = np.where(np.abs(x) <= a_el, x, (a_el * np.sign(x)).astype(F)).astype(F)
            q = qo.mx_div(np.broadcast_to(np.asarray(a_el, F), x.shape), F(255))
            want, _ = qo.qdq(c, q)
            ok = bits(alpha.cpu().numpy(), alpha_ref) and bits(yd.cpu().numpy(), want)
            if not ok:
                failures += 1
                print("rank %d FAIL GDRQ_PY %s group_size=%d step=%d alpha=%r want=%r" % (
                    rank, mode, group_size, step, alpha.cpu().numpy(), alpha_ref), flush=True)
        if ex is not None:
            torch.cuda.synchronize()
            dist.barrier()
            ex.close()

    # peer-memory allreduce (gradients: sum / average in rank order; statistic vectors: max) against the same reduction
    # done on the host from the all-gathered inputs -- bit for bit, identical on every rank
    from b200quant.dist import PeerBuffer, PeerGradBucket, PeerThresholdExchange
    ex = PeerThresholdExchange(torch.device("cuda", local))
    big = PeerBuffer(ex, (1 << 22) + 3)
    for it, count in enumerate([1, 3, 4, 5, 8, 1023, 4096, 65537, (1 << 20) + 1, (1 << 22) + 3]):
        for kind in ("sum", "avg", "max"):
            rng = np.random.default_rng(31 * it + 7 * rank + len(kind))
            mine = (rng.standard_normal(count) * (1 + rank)).astype(F)
            if kind == "max" and count > 4:
                mine[rank % count] = np.nan if rank == world - 1 else mine[rank % count]      # NaN propagates
            big.tensor[:count].copy_(torch.from_numpy(mine))
            big.tensor[count:count + 2] = 7.0 if count + 2 <= big.numel else 0.0
            everyone = [torch.zeros(count, device="cuda") for _ in range(world)]
            dist.all_gather(everyone, torch.from_numpy(mine).cuda())
            if kind == "max":
                big.allreduce_max(count)
            else:
                big.allreduce_sum(count, average=(kind == "avg"))
            got = big.tensor[:count].cpu().numpy()
            acc = everyone[0].cpu().numpy().copy()
            for r in range(1, world):
                o = everyone[r].cpu().numpy()
                acc = np.where(np.isnan(acc) | np.isnan(o), F(np.nan), np.maximum(acc, o)).astype(F) if kind == "max" else (acc + o).astype(F)
            if kind == "avg":
                acc = (acc * F(1.0 / world)).astype(F)
            ok = bits(got, acc) if kind != "max" else (np.array_equal(np.isnan(got), np.isnan(acc)) and
                                                       np.array_equal(got[~np.isnan(got)], acc[~np.isnan(acc)]))
            if count + 2 <= big.numel and not bool((big.tensor[count:count + 2] == 7.0).all()):
                ok = False                                      # nothing beyond `count` may be touched
            if not ok:
                failures += 1
                print("rank %d FAIL peer allreduce %s count=%d" % (rank, kind, count), flush=True)
    shapes = [(64, 3, 7, 7), (64, 64, 1, 1), (512, 512, 3, 3), (1000, 2048), (7,)]
    bucket = PeerGradBucket(shapes, ex)
    for step in range(3):
        rng = np.random.default_rng(900 + step * 10 + rank)
        for v in bucket.views:
            v.copy_(torch.from_numpy(rng.standard_normal(tuple(v.shape)).astype(F)))
        ref = bucket.flat.clone()
        dist.all_reduce(ref, op=dist.ReduceOp.SUM)
        ref /= world
        bucket.allreduce(average=True)
        if not torch.allclose(bucket.flat, ref, rtol=1e-6, atol=1e-6):
            failures += 1
            print("rank %d FAIL PeerGradBucket step %d" % (rank, step), flush=True)
        gathered = [torch.zeros_like(bucket.flat) for _ in range(world)]
        dist.all_gather(gathered, bucket.flat)
        if not all(torch.equal(gathered[0].view(torch.int32), g.view(torch.int32)) for g in gathered):
            failures += 1
            print("rank %d FAIL PeerGradBucket not identical across ranks" % rank, flush=True)
    torch.cuda.synchronize()
    dist.barrier()
    if ex.status() is not None:
        failures += 1
        print("rank %d FAIL peer exchange timed out %r" % (rank, ex.status()), flush=True)
    bucket.close()
    big.close()
    ex.close()

    t = torch.tensor([failures], device="cuda")
    dist.all_reduce(t)
    if rank == 0:
        print("multi_gpu_check: world=%d failures=%d" % (world, int(t.item())), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 1 if int(t.item()) else 0


if __name__ == "__main__":
    sys.exit(main())
