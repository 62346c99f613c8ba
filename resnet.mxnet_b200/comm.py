"""One process driving N devices -- the reference's own layout (``train.py:34`` builds ``[mx.gpu(i) ...]``,
``core/solver.py:58-61`` binds one Module over them, the KVStore sums gradients, ``solver.py:121``) -- without
``torch.distributed``, NCCL or CUDA IPC: ``b2q_comm_*`` of include/b2q.h.

    group = DeviceGroup([0, 1, 2, 3])
    group.attach_threshold_exchange(ops_per_rank)      # activation nodes: fused peer-memory exchange, as in dist.py
    group.allreduce_sum(grads_per_rank, average=True)   # weight gradients: slice-owner kernels over NVLink, in place
    group.allreduce_max(stats_per_rank)

Arrays are one per rank, each on its rank's device; the collectives are asynchronous and ordered with the arrays'
current streams on every device.
"""
import ctypes

from . import _lib
from .dlpack import as_buffer, current_stream


class DeviceGroup(object):
    def __init__(self, devices):
        self.devices = [int(d) for d in devices]
        if len(set(self.devices)) != len(self.devices):
            raise ValueError("one rank per device")
        self.world = len(self.devices)
        self.ctxs = [_lib.context(d) for d in self.devices]
        self.lib = _lib.load()
        handles = (ctypes.c_void_p * self.world)(*[c.handle.value for c in self.ctxs])
        self.handle = ctypes.c_void_p()
        rc = self.lib.b2q_comm_create(handles, self.world, ctypes.byref(self.handle))
        if rc != 0:
            raise _lib.B2QError("b2q_comm_create failed (%d): %s" % (rc, _lib.last_error()))
        self._exchanges = []

    def _check(self, rc, what):
        if rc != 0:
            raise _lib.B2QError("%s failed (%d): %s" % (what, rc, _lib.last_error()))

    def mailboxes(self, rank):
        boxes = ctypes.POINTER(ctypes.c_void_p)()
        self._check(self.lib.b2q_comm_mailboxes(self.handle, int(rank), ctypes.byref(boxes)), "b2q_comm_mailboxes")
        return [boxes[r] for r in range(self.world)]

    def exchange(self, rank):
        """dist.PeerThresholdExchange of ``rank`` (what ``op.peer`` expects), backed by the communicator's mailboxes."""
        from .dist import PeerThresholdExchange
        ex = PeerThresholdExchange.from_tables(self.ctxs[rank], self.mailboxes(rank), rank, self.world)
        self._exchanges.append(ex)
        return ex

    def attach_threshold_exchange(self, ops_per_rank):
        """ops_per_rank[r]: the operators of rank r's replica, in network order (every rank the same sequence)."""
        assert len(ops_per_rank) == self.world
        out = []
        for r, ops in enumerate(ops_per_rank):
            ex = self.exchange(r)
            for op in ops:
                kind = op.__class__.__name__
                if (not getattr(op, "is_weight", True) and hasattr(op, "VARIANT")) \
                        or (kind == "GDRQ_PY" and not op.is_weight and op.group_size == -1) or kind == "GDRQ_Fold_BN":
                    op.peer, op.sync = ex, None
            out.append(ex)
        return out

    def _collect(self, arrays):
        assert len(arrays) == self.world
        bufs = [as_buffer(a, write=True) for a in arrays]
        n = bufs[0].numel
        for r, b in enumerate(bufs):
            if not b.on_device or b.device_id != self.devices[r]:
                raise ValueError("array %d must live on cuda:%d" % (r, self.devices[r]))
            if b.numel != n:
                raise ValueError("all ranks must pass the same number of elements")
        ptrs = (ctypes.c_void_p * self.world)(*[b.ptr for b in bufs])
        streams = (ctypes.c_void_p * self.world)(*[current_stream(b) for b in bufs])
        return ptrs, streams, n, bufs

    def allreduce_max(self, arrays):
        ptrs, streams, n, keep = self._collect(arrays)
        self._check(self.lib.b2q_comm_allreduce_max_f32(self.handle, ptrs, n, streams), "b2q_comm_allreduce_max_f32")
        return arrays

    def allreduce_sum(self, arrays, average=False):
        ptrs, streams, n, keep = self._collect(arrays)
        self._check(self.lib.b2q_comm_allreduce_sum_f32(self.handle, ptrs, n, int(bool(average)), streams),
                    "b2q_comm_allreduce_sum_f32")
        return arrays

    def check(self):
        for ex in self._exchanges:
            ex.check()

    def close(self):
        if self.handle:
            self.lib.b2q_comm_destroy(self.handle)
            self.handle = ctypes.c_void_p()
