"""torch driver for the CustomOp classes (the only runnable host in an image without MXNet).

``Custom`` plays the role of ``mx.sym.Custom(..., op_type=...)`` plus MXNet's executor for one node: it builds
the operator through its registered Prop from *string* attributes, owns the auxiliary states as buffers named
like the reference's (``minmax``, ``alpha``, ``data_minmax`` ...), and calls ``forward`` / ``backward`` with
the MXNet protocol (lists of arrays, ``req`` strings, ``is_train``).  torch only supplies device memory, the
stream and autograd bookkeeping; all arithmetic is in libb2q.so.
"""
import torch
import torch.nn as nn

from .operator import get_prop


class _CustomFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, node, *inputs):
        in_data = [t.detach().contiguous() for t in inputs]
        out = node._alloc_out(in_data)
        node.op.forward(bool(node.training), ["write"], in_data, [out], node.aux_list())
        ctx.node = node
        ctx.in_data = in_data
        ctx.out = out
        ctx.needs = [bool(t.requires_grad) for t in inputs]
        return out

    @staticmethod
    def backward(ctx, grad_out):
        node = ctx.node
        in_data = ctx.in_data
        if node.alias_ste and node.is_identity_backward():
            return (None, grad_out) + (None,) * (len(in_data) - 1)
        grads = [torch.empty_like(t) for t in in_data]
        node.op.backward(["write"] * len(in_data), [grad_out.contiguous()], in_data, [ctx.out], grads,
                         node.aux_list())
        return (None,) + tuple(g if need else None for g, need in zip(grads, ctx.needs))


class Custom(nn.Module):
    """One quantization node.  ``Custom("Quantization_int8_V2", is_weight=True, quant_mode="minmax")(w)``."""

    def __init__(self, op_type, aux_init=1.0, alias_ste=False, **attrs):
        super(Custom, self).__init__()
        self.op_type = op_type
        self.attrs = {k: str(v) for k, v in attrs.items()}   # MXNet hands attributes over as strings
        self.prop = get_prop(op_type)(**self.attrs)
        self.op = self.prop.create_operator(None, None, None)
        self.aux_names = list(self.prop.list_auxiliary_states())
        self.aux_init = aux_init
        self.alias_ste = alias_ste
        self._aux_ready = False

    # -- auxiliary states ------------------------------------------------------------------------------
    def _ensure_aux(self, in_data):
        if self._aux_ready:
            return
        _, _, aux_shapes = self.prop.infer_shape([list(t.shape) for t in in_data])
        init = self.aux_init if isinstance(self.aux_init, (list, tuple)) else [self.aux_init] * len(self.aux_names)
        for name, shape, v in zip(self.aux_names, aux_shapes, init):
            if getattr(self, name, None) is None:
                self.register_buffer(name, torch.full(tuple(shape), float(v), dtype=torch.float32,
                                                      device=in_data[0].device))
        self._aux_ready = True

    def aux_list(self):
        return [getattr(self, n) for n in self.aux_names]

    def _alloc_out(self, in_data):
        self._ensure_aux(in_data)
        _, out_shapes, _ = self.prop.infer_shape([list(t.shape) for t in in_data])
        return torch.empty(tuple(out_shapes[0]), dtype=torch.float32, device=in_data[0].device)

    def is_identity_backward(self):
        op = self.op
        if self.op_type in ("Quantization_int8_V2", "QUANT_STE_PY"):
            return True
        if self.op_type in ("ClipGrad_Quantization_int8", "GDRQ_PY"):
            return bool(op.is_weight)
        return False

    # -- op-instance state the reference forgets to checkpoint ---------------------
    def get_extra_state(self):
        return {k: getattr(self.op, k) for k in ("delay_quant", "init") if hasattr(self.op, k)}

    def set_extra_state(self, state):
        for k, v in (state or {}).items():
            setattr(self.op, k, v)

    def forward(self, *inputs):
        return _CustomFn.apply(self, *inputs)

    def extra_repr(self):
        return "op_type=%s, %s" % (self.op_type, ", ".join("%s=%s" % kv for kv in sorted(self.attrs.items())))
