"""Drop-in for ``core/operator/GDRQ.py``: op_types ``GDRQ_PY`` and ``CLIP_RELU_PY``.

GDRQ_PY (GDRQ.py:50-152): threshold alpha from ktimes*mean|x| (weights: alpha := thr; activations:
alpha += lamda*(alpha - thr), also in eval mode, as in the reference), clip to +-alpha, round to 2^nbits-1
levels unless still in the delay_quant countdown.  Grouped mode indexes the NCHW / OIHW tensor directly as
(outer, groups, inner) instead of the reference's swapaxes -> reshape -> ... -> swapaxes round trip (:88-118).
"""
from .. import _kernels as K
from .. import _lib
from ..operator import CustomOp, CustomOpProp, py_bool, py_literal, register


class GDRQ_PY(CustomOp):
    def __init__(self, nbits, group_size, is_weight, lamda, delay_quant, fix_alpha, ktimes):
        self.nbits = nbits
        self.group_size = group_size
        self.is_weight = is_weight
        self.lamda = lamda
        self.delay_quant = delay_quant
        self.QUANT_LEVEL = 2 ** (self.nbits) - 1
        self.fix_alpha = fix_alpha
        self.ktimes = ktimes
        self.sync = None        # dist.ThresholdSync: max over ranks of the mean|x| statistic (activations only)
        self.peer = None        # dist.PeerThresholdExchange: the same exchange fused into the two forward kernels
        self._stat = None

    def forward(self, is_train, req, in_data, out_data, aux):
        do_round = not (self.delay_quant > 0)          # GDRQ.py:81-85 / :110-114
        if self.delay_quant > 0:
            self.delay_quant -= 1
        if (self.peer is not None and not self.is_weight and not self.fix_alpha and do_round and self.group_size == -1
                and req[0] in ("write", "inplace")):
            self.peer.quantize_mean(_lib.UPD_GDRQ_ACT, in_data[0], out_data[0], aux[0], self.ktimes, self.lamda,
                                    self.QUANT_LEVEL)
            return
        if (self.sync is not None or self.peer is not None) and not self.is_weight and not self.fix_alpha:
            # data parallel: mean|x| per group -> allreduce(max) -> alpha update -> clip + round, so that every rank
            # applies the same alpha (k*max(mean) == max(k*mean): the scaling is monotonic)
            x, alpha = in_data[0], aux[0]
            view = K._gdrq_view(tuple(x.shape), self.group_size, False)
            if self.sync is None and getattr(self.peer, "vectors", False):
                # per-group statistics in peer-addressable memory, maximised over ranks by one kernel per rank
                stat = self.peer.stat_vector(view[1])
                K.meanabs(x, stat, view)
                self.peer.allreduce_max_vector(view[1])
            else:
                if self._stat is None:
                    self._stat = K.scratch_like(x, view[1])
                stat = self._stat
                K.meanabs(x, stat, view)
                if self.sync is None:   # an exchange without vectors (comm.DeviceGroup): NCCL for this call
                    from ..dist import ThresholdSync
                    self.sync = ThresholdSync()
                self.sync(stat)
            K.threshold_update(_lib.UPD_GDRQ_ACT, stat, alpha, self.ktimes, self.lamda)
            K.qdq(x, out_data[0], alpha, self.QUANT_LEVEL, _lib.CLIP_SYM if view[1] == 1 else _lib.CLIP_WHERE_LE,
                  req[0], view=view, do_round=do_round)
            return
        K.gdrq_fwd(in_data[0], out_data[0], aux[0], self.group_size, self.is_weight, self.fix_alpha, do_round,
                   self.QUANT_LEVEL, self.ktimes, self.lamda, req[0])

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        if self.is_weight:
            K.ste_bwd(out_grad[0], in_grad[0], req[0])  # :125-127
        else:
            K.gdrq_bwd(in_data[0], out_grad[0], in_grad[0], aux[0], self.group_size, req[0])   # :131-152


@register("GDRQ_PY")
class GDRQ_PYProp(CustomOpProp):
    def __init__(self, nbits=4, group_size=-1, is_weight=False, lamda=0.001, delay_quant=0, fix_alpha=False, ktimes=3):
        self.nbits = int(nbits)
        self.group_size = int(group_size)
        self.is_weight = py_bool(is_weight)
        self.lamda = float(lamda)
        self.delay_quant = int(delay_quant)
        self.fix_alpha = py_bool(fix_alpha)
        self.ktimes = float(ktimes)
        super(GDRQ_PYProp, self).__init__(True)

    def list_arguments(self):
        return ["data"]

    def list_outputs(self):
        return ["output"]

    def list_auxiliary_states(self):
        return ["alpha"]

    def infer_shape(self, in_shape):
        shape = in_shape[0]
        if self.group_size == -1:
            aux_shape = [1]
        else:
            channels = shape[0] if self.is_weight else shape[1]
            assert channels % self.group_size == 0, \
                "the channels of weight or activation must be divisible by group size. channels({}) vs group size({})." \
                .format(channels, self.group_size)
            aux_shape = [channels // self.group_size]
        return [shape], [shape], [aux_shape]

    def create_operator(self, ctx, shapes, dtypes):
        return GDRQ_PY(self.nbits, self.group_size, self.is_weight, self.lamda, self.delay_quant, self.fix_alpha,
                       self.ktimes)


class CLIP_RELU_PY(CustomOp):
    """GDRQ.py:192-208."""

    def __init__(self, nbits, threshold):
        self.nbits = nbits
        self.threshold = threshold
        self.QUANT_LEVEL = 2 ** (self.nbits) - 1
        self.count = 0

    def forward(self, is_train, req, in_data, out_data, aux):
        K.clip_relu_fwd(in_data[0], out_data[0], self.threshold, self.QUANT_LEVEL, req[0])

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        K.mask_bwd(in_data[0], out_grad[0], in_grad[0], None, self.threshold, _lib.MASK_LT, req[0])


@register("CLIP_RELU_PY")
class CLIP_RELU_PYProp(CustomOpProp):
    def __init__(self, nbits="8", threshold="8.0"):
        self.nbits = py_literal(nbits)
        self.threshold = py_literal(threshold)
        super(CLIP_RELU_PYProp, self).__init__(True)

    def infer_shape(self, in_shape):
        return [in_shape[0]], [in_shape[0]], []

    def create_operator(self, ctx, shapes, dtypes):
        return CLIP_RELU_PY(self.nbits, self.threshold)
