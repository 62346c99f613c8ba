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
            if average and self.flat.is_cuda and dist.get_backend(group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)   # the division rides in the collective
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                if average:
                    self.flat.div_(dist.get_world_size(group))
        return self.flat


class PeerThresholdExchange(object):
    """Threshold exchange over peer memory, fused into the forward kernels (b2q_peer_minmax_quant_fwd_f32): no NCCL
    call on the forward critical path.  Each rank allocates a mailbox, the 64-byte CUDA IPC handles are exchanged
    once through torch.distributed, and every activation node then runs [reduce + publish] -> [poll + update + QDQ].
    Single node only (CUDA IPC); all ranks must call the nodes in the same order."""

    def __init__(self, device, group=None):
        import ctypes
        from . import _lib
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.device(device)
        self.ctx = _lib.context(self.device.index)
        self.sequence = 0
        self.vectors = True      # stat_vector / allreduce_max_vector available (needs the process group for the handles)
        self._stats = None
        own = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        self.ctx.call("b2q_peer_mailbox_create", ctypes.byref(own), ctypes.cast(handle, ctypes.c_void_p))
        self._own = own
        handles = [None] * self.world
        if self.world > 1:
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
        else:
            handles[0] = bytes(handle.raw)
        self._boxes = (ctypes.c_void_p * self.world)()
        self._opened = []
        for r in range(self.world):
            if r == self.rank:
                self._boxes[r] = own.value
            else:
                p = ctypes.c_void_p()
                buf = ctypes.create_string_buffer(handles[r], 64)
                self.ctx.call("b2q_peer_mailbox_open", ctypes.cast(buf, ctypes.c_void_p), ctypes.byref(p))
                self._boxes[r] = p.value
                self._opened.append(p)
        if self.world > 1:
            dist.barrier(group=group)

    @classmethod
    def from_tables(cls, ctx, boxes, rank, world, own=None):
        """Exchange over mailboxes that already exist and are mapped (b200quant.comm.DeviceGroup: one process, N devices)."""
        import ctypes
        self = cls.__new__(cls)
        self.group, self.world, self.rank, self.ctx, self.sequence = None, int(world), int(rank), ctx, 0
        self.device = torch.device("cuda", ctx.device)
        self._boxes = (ctypes.c_void_p * self.world)(*[boxes[r] for r in range(self.world)])
        self._opened, self._own = [], None      # the mailboxes belong to the communicator
        self.vectors, self._stats = False, None
        self._status_box = ctypes.c_void_p(boxes[rank]) if own is None else own
        return self

    def quantize(self, variant, x, y, aux, init, ema_decay):
        """Activation forward (training) of Quantization_int8_V2 (variant 0) / ClipGrad (variant 1) with the
        statistic maximised over all ranks."""
        import numpy as np
        from .dlpack import as_buffer, current_stream
        xb, yb, ab = as_buffer(x), as_buffer(y, write=True), as_buffer(aux, write=True)
        self.sequence += 1
        self.ctx.call("b2q_peer_minmax_quant_fwd_f32", int(variant), xb.ptr, yb.ptr, ab.ptr, xb.numel, int(bool(init)),
                      float(np.float32(ema_decay)), float(np.float32(1 - ema_decay)), self._boxes, self.rank, self.world,
                      self.sequence, current_stream(xb))

    def quantize_mean(self, upd_mode, x, y, aux, p0, p1, qlevel):
        """Whole-tensor activation forward of the mean-based operators (GDRQ_PY: upd_mode UPD_GDRQ_ACT, p0 = ktimes,
        p1 = lamda; GDRQ_Fold_BN data: UPD_TWICE_STORE / UPD_TWICE_EMA, p0 = ema_decay, p1 = 1 - ema_decay) with the
        mean|x| statistic maximised over all ranks."""
        import numpy as np
        from .dlpack import as_buffer, current_stream
        xb, yb, ab = as_buffer(x), as_buffer(y, write=True), as_buffer(aux, write=True)
        self.sequence += 1
        self.ctx.call("b2q_peer_meanabs_quant_fwd_f32", int(upd_mode), xb.ptr, yb.ptr, ab.ptr, xb.numel,
                      float(np.float32(p0)), float(np.float32(p1)), float(np.float32(qlevel)), self._boxes, self.rank,
                      self.world, current_stream(xb))

    def stat_vector(self, n):
        """A float32 vector of ``n`` elements in peer-addressable memory for per-group statistics (grouped GDRQ_PY
        activations); one buffer per exchange, shared by its operators (stream order makes that safe: the allreduce's
        closing barrier means no rank still reads it when the next operator writes)."""
        if getattr(self, "_stats", None) is None:
            self._stats = PeerBuffer(self, 8192)
        if n > self._stats.numel:
            raise ValueError("more than %d groups" % self._stats.numel)
        return self._stats.tensor[:n]

    def allreduce_max_vector(self, n):
        """max over ranks of the first ``n`` elements of ``stat_vector`` (one kernel per rank, no NCCL)."""
        self._stats.allreduce_max(n)

    def status(self):
        """(sequence, rank) of the first exchange in which this rank gave up waiting for a peer's statistic (option
        ``peer_timeout_ms``; the affected outputs and thresholds are NaN), or None.  Synchronises with the device."""
        import ctypes
        seq, rk = ctypes.c_uint32(0), ctypes.c_uint32(0)
        box = self._own if self._own is not None else getattr(self, "_status_box", None)
        if box is None:
            return None
        self.ctx.call("b2q_peer_status", box, ctypes.byref(seq), ctypes.byref(rk))
        return (int(seq.value), int(rk.value)) if seq.value else None

    def check(self):
        st = self.status()
        if st is not None:
            from . import _lib
            raise _lib.B2QError("rank %d waited longer than peer_timeout_ms for the statistic of rank %d in exchange %d; "
                                "thresholds and outputs from that call on are NaN" % (self.rank, st[1], st[0]))

    def close(self):
        if self._own is not None:
            try:
                self.check()
            finally:
                self._close()

    def _close(self):
        if getattr(self, "_stats", None) is not None:
            self._stats.close()
            self._stats = None
        for p in self._opened:
            self.ctx.call("b2q_peer_mailbox_close", p)
        self._opened = []
        if self._own is not None:
            self.ctx.call("b2q_peer_mailbox_destroy", self._own)
            self._own = None


def pin_rank_to_host_cores(local_rank, local_world, device_index=None):
    """Give every rank of a node its own host cores (and with them, by first touch, its own host memory): the cores of the
    GPU's NUMA node when the platform reports one (``/sys/bus/pci/devices/<bus id>/numa_node``), otherwise an even split
    of the cores this process may use.  The host-buffer path (pinned staging, the host-to-host straight-through copy pool)
    then neither migrates between sockets nor fights the other ranks for the same cores.  Returns the core list; sets
    ``B2Q_HOST_COPY_THREADS`` to its length unless the variable is already set.  Call before the first host-buffer call."""
    import os
    try:
        allowed = sorted(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return []
    cores = None

    def numa_node_of(dev):
        prop = torch.cuda.get_device_properties(dev)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id)
        return int(open(path).read().strip())

    if device_index is not None and torch.cuda.is_available():
        try:
            node = numa_node_of(device_index)
            if node >= 0:
                spans = open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(",")
                on_node = []
                for sp in spans:
                    lo, _, hi = sp.partition("-")
                    on_node += list(range(int(lo), int(hi or lo) + 1))
                on_node = [c for c in on_node if c in allowed]
                # the local ranks whose GPU hangs off the same node share its cores evenly, in rank order (local rank r
                # drives device r, as torchrun sets it up)
                same = []
                for r in range(local_world):
                    try:
                        if numa_node_of(r) == node:
                            same.append(r)
                    except Exception:
                        pass
                if local_rank not in same:
                    same.append(local_rank)
                    same.sort()
                share = max(1, len(on_node) // len(same))
                k = same.index(local_rank)
                if on_node:
                    cores = on_node[k * share:(k + 1) * share] or on_node
        except Exception:
            cores = None
    if cores is None:
        share = max(1, len(allowed) // max(1, local_world))
        cores = allowed[local_rank * share:(local_rank + 1) * share] or allowed
    try:
        os.sched_setaffinity(0, cores)
    except OSError:  # pragma: no cover
        return allowed
    os.environ.setdefault("B2Q_HOST_COPY_THREADS", str(max(1, len(cores))))
    return cores


class _RawCudaArray(object):
    """Lets torch wrap device memory this package allocated (``torch.as_tensor`` reads ``__cuda_array_interface__``)."""

    def __init__(self, ptr, numel):
        self.__cuda_array_interface__ = {"shape": (int(numel),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class PeerBuffer(object):
    """A float32 device buffer every rank of the exchange can address: allocated by the library (cudaMalloc + CUDA IPC
    handle), the handles all-gathered once through torch.distributed, the peers' buffers mapped into this process.
    ``tensor`` is the rank's own buffer as a torch tensor (zero-filled); ``allreduce_sum`` / ``allreduce_max`` reduce the
    first ``count`` elements over all ranks in place with ONE kernel per rank over NVLink (b2q_peer_allreduce_*_f32): no
    NCCL call.  All ranks must issue the same calls in the same order, and the allreduce calls of one exchange must be
    ordered with each other on the device (one stream, or streams that wait for each other): they share the exchange's
    sequence counter.  ``tensor`` is a view of library-owned memory: do not use it after ``close()``."""

    def __init__(self, exchange, numel):
        import ctypes
        self.ex = exchange
        self.ctx = exchange.ctx
        self.numel = int(numel)
        own = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(64)
        self.ctx.call("b2q_peer_buffer_create", 4 * max(self.numel, 1), ctypes.byref(own), ctypes.cast(handle, ctypes.c_void_p))
        self._own = own
        world, rank = exchange.world, exchange.rank
        handles = [None] * world
        if world > 1:
            dist.all_gather_object(handles, bytes(handle.raw), group=exchange.group)
        else:
            handles[0] = bytes(handle.raw)
        self._ptrs = (ctypes.c_void_p * world)()
        self._opened = []
        for r in range(world):
            if r == rank:
                self._ptrs[r] = own.value
            else:
                p = ctypes.c_void_p()
                buf = ctypes.create_string_buffer(handles[r], 64)
                self.ctx.call("b2q_peer_mailbox_open", ctypes.cast(buf, ctypes.c_void_p), ctypes.byref(p))
                self._ptrs[r] = p.value
                self._opened.append(p)
        self.tensor = torch.as_tensor(_RawCudaArray(own.value, max(self.numel, 1)), device=exchange.device)[:self.numel]
        if world > 1:
            dist.barrier(group=exchange.group)

    def _stream(self):
        return torch.cuda.current_stream(self.ex.device).cuda_stream

    def allreduce_sum(self, count=None, average=False):
        n = self.numel if count is None else int(count)
        self.ctx.call("b2q_peer_allreduce_sum_f32", self._ptrs, n, int(bool(average)), self.ex._boxes, self.ex.rank,
                      self.ex.world, self._stream())
        return self.tensor

    def allreduce_max(self, count=None):
        n = self.numel if count is None else int(count)
        self.ctx.call("b2q_peer_allreduce_max_f32", self._ptrs, n, self.ex._boxes, self.ex.rank, self.ex.world,
                      self._stream())
        return self.tensor

    def close(self):
        self.tensor = None
        for p in self._opened:
            self.ctx.call("b2q_peer_mailbox_close", p)
        self._opened = []
        if self._own is not None:
            self.ctx.call("b2q_peer_mailbox_destroy", self._own)
            self._own = None


class PeerGradBucket(object):
    """GradBucket whose flat buffer lives in peer-addressable memory: the weight-node ``in_grad`` views are slices of it
    and ``allreduce`` is one slice-owner kernel per rank over NVLink instead of an NCCL call (KVStore 'device' semantics,
    core/solver.py:121: summed -- averaged with ``average=True`` -- in rank order, bit-identical on every rank)."""

    def __init__(self, shapes, exchange):
        self.numel = [int(torch.Size(s).numel()) for s in shapes]
        self.buffer = PeerBuffer(exchange, sum(self.numel))
        self.flat = self.buffer.tensor
        self.views = []
        off = 0
        for s, n in zip(shapes, self.numel):
            self.views.append(self.flat[off:off + n].view(s))
            off += n

    def allreduce(self, group=None, average=True):
        return self.buffer.allreduce_sum(average=average)

    def close(self):
        self.views, self.flat = [], None
        self.buffer.close()


def attach_peer_exchange(ops, device, group=None):
    """Route every activation node in ``ops`` that has a fused peer-memory exchange through it: the minmax operators,
    GDRQ_PY activations and the data path of GDRQ_Fold_BN.  Grouped GDRQ activations (one threshold per channel
    group) exchange their statistic vector with the peer-memory allreduce(max) kernel between the reduction and the
    threshold update."""
    ex = PeerThresholdExchange(device, group)
    for op in ops:
        kind = op.__class__.__name__
        if (not getattr(op, "is_weight", True) and hasattr(op, "VARIANT")) \
                or (kind == "GDRQ_PY" and not op.is_weight) or kind == "GDRQ_Fold_BN":
            op.peer = ex       # grouped GDRQ_PY activations: per-group statistics through ex.allreduce_max_vector
            op.sync = None
    return ex
