"""Multi-tensor execution of the weight-quantization nodes: all weights of a network in two launches.

    group = WeightGroup(ops, weights, outputs, auxs, out_grads, in_grads)
    group.forward(is_train)     # == op.forward(is_train, ['write'], [w], [y], [aux]) for every op
    group.backward()            # == op.backward(['write'], [dy], ..., [dx], ...) for every op

The operators stay the source of truth for attributes and state: while any of them is still in its ``delay_quant``
countdown the group falls back to the per-operator calls, so results are identical to driving the ops one by one
(tests/test_gpu_multi.py).
"""
import ctypes

from . import _lib
from .dlpack import as_buffer, current_stream


class WeightGroup(object):
    def __init__(self, ops, weights, outputs, auxs, out_grads=None, in_grads=None):
        assert len(ops) == len(weights) == len(outputs) == len(auxs) and len(ops) > 0
        kinds = {type(op).__name__ for op in ops}
        if len(kinds) != 1 or not all(op.is_weight for op in ops):
            raise ValueError("WeightGroup takes weight nodes of one operator type")
        self.kind = kinds.pop()
        if self.kind not in ("Quantization_int8", "ClipGrad_Quantization_int8", "GDRQ_PY"):
            raise ValueError("WeightGroup supports Quantization_int8_V2, ClipGrad_Quantization_int8 and GDRQ_PY weights")
        self.ops, self.weights, self.outputs, self.auxs = list(ops), list(weights), list(outputs), list(auxs)
        self.out_grads = list(out_grads) if out_grads is not None else None
        self.in_grads = list(in_grads) if in_grads is not None else None
        self.gdrq = self.kind == "GDRQ_PY"
        if self.gdrq:
            sig = {(op.nbits, op.group_size if op.group_size == -1 else "g", op.fix_alpha, op.ktimes) for op in ops}
            if len({(s[0], s[2], s[3]) for s in sig}) != 1:
                raise ValueError("GDRQ weight nodes of one WeightGroup must share nbits / fix_alpha / ktimes")
        self.variant = None if self.gdrq else ops[0].VARIANT
        descs = (_lib.WeightDesc * len(ops))()
        self._keep = []
        dev = None
        for i, op in enumerate(ops):
            xb, yb, ab = as_buffer(weights[i]), as_buffer(outputs[i], write=True), as_buffer(auxs[i], write=True)
            if not (xb.on_device and yb.on_device and ab.on_device):
                raise _lib.B2QError("WeightGroup needs CUDA tensors")
            dev = xb.device_id if dev is None else dev
            if xb.device_id != dev:
                raise ValueError("all tensors of a WeightGroup must live on one device")
            rows = int(xb.shape[0])
            cols = xb.numel // rows
            if self.gdrq:   # grouped weights: one "row" per group of group_size channels (GDRQ.py:92-94)
                pc = op.group_size != -1
                if pc:
                    if rows % op.group_size:
                        raise ValueError("channels must be divisible by group_size")
                    rows, cols = rows // op.group_size, cols * op.group_size
            else:
                pc = bool(op.is_weight_perchannel)
            if ab.numel != (rows if pc else 1):
                raise ValueError("aux %d has %d elements" % (i, ab.numel))
            d = descs[i]
            d.x, d.y, d.aux, d.rows, d.cols, d.per_channel = xb.ptr, yb.ptr, ab.ptr, rows, cols, int(pc)
            if self.out_grads is not None and self.in_grads is not None:
                gb, ib = as_buffer(self.out_grads[i]), as_buffer(self.in_grads[i], write=True)
                d.dy, d.dx = gb.ptr, ib.ptr
                self._keep += [gb, ib]
            self._keep += [xb, yb, ab]
        self.device = dev
        self.ctx = _lib.context(dev)
        self._stream_of = as_buffer(weights[0])
        self.plan = ctypes.c_void_p()
        self.ctx.call("b2q_multi_plan_create", ctypes.cast(descs, ctypes.c_void_p), len(ops), ctypes.byref(self.plan))

    def _delayed(self, is_train):
        return bool(is_train) and any(op.delay_quant > 0 for op in self.ops)

    def _per_op(self, is_train):
        for op, w, y, a in zip(self.ops, self.weights, self.outputs, self.auxs):
            op.forward(is_train, ["write"], [w], [y], [a])

    def forward(self, is_train):
        if self.gdrq:
            delays = {op.delay_quant > 0 for op in self.ops}
            if len(delays) != 1:   # mixed countdowns: let every operator handle its own state
                return self._per_op(is_train)
            do_round = not delays.pop()        # GDRQ.py:81-85: clip only while delay_quant counts down
            if not do_round:
                for op in self.ops:
                    op.delay_quant -= 1
            op0 = self.ops[0]
            import numpy as np
            self.ctx.call("b2q_multi_gdrq_weight_fwd_f32", self.plan, int(bool(op0.fix_alpha)), int(do_round),
                          float(np.float32(op0.QUANT_LEVEL)), float(np.float32(op0.ktimes)),
                          current_stream(self._stream_of))
            return
        if self._delayed(is_train):
            return self._per_op(is_train)
        self.ctx.call("b2q_multi_weight_quant_fwd_f32", self.plan, int(self.variant), int(bool(is_train)),
                      current_stream(self._stream_of))

    def backward(self):
        if self.out_grads is None or self.in_grads is None:
            raise ValueError("WeightGroup was built without gradient tensors")
        self.ctx.call("b2q_multi_weight_ste_bwd_f32", self.plan, current_stream(self._stream_of))

    def close(self):
        if self.plan:
            self.ctx.call("b2q_multi_plan_destroy", self.plan)
            self.plan = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
