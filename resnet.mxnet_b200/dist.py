"""Data-parallel plumbing for the quantization nodes: one process per GPU, torch.distributed (NCCL over
NVLink / NVSwitch; gloo on CPU for tests).

Two exchanges per training step:
  * thresholds -- allreduce(MAX) of the per-rank activation statistic *before* the EMA / alpha update, so every
    rank holds identical aux.  (The reference keeps per-device aux and only averages them once per epoch,
    core/solver.py:170-171; single-rank results are unaffected.)
  * gradients  -- allreduce(SUM) of the weight gradients, bucketed into one flat buffer.
"""
import torch
import torch.distributed as dist


class ThresholdSync(object):
    """Callable handed to an operator (``op.sync = ThresholdSync()``): reduces its statistic across ranks."""

    def __init__(self, group=None):
        self.group = group
        self.calls = 0

    def __call__(self, stat):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(stat, op=dist.ReduceOp.MAX, group=self.group)
        self.calls += 1
        return stat


def attach_threshold_sync(ops, group=None):
    """Give every activation-quantization operator in ``ops`` a cross-rank threshold sync."""
    sync = ThresholdSync(group)
    for op in ops:
        if not getattr(op, "is_weight", True) or op.__class__.__name__ == "GDRQ_Fold_BN":
            op.sync = sync
    return sync


class GradBucket(object):
    """One flat float32 buffer holding all weight gradients so a single allreduce(SUM) covers them."""

    def __init__(self, shapes, device):
        self.numel = [int(torch.Size(s).numel()) for s in shapes]
        self.flat = torch.zeros(sum(self.numel), dtype=torch.float32, device=device)
        self.views = []
        off = 0
        for s, n in zip(shapes, self.numel):
            self.views.append(self.flat[off:off + n].view(s))
            off += n

    def allreduce(self, group=None, average=True):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                self.flat.div_(dist.get_world_size(group))
        return self.flat
