"""ctypes binding of libb2q.so (C ABI declared in include/b2q.h).

There is deliberately no fallback: if the CUDA library is missing or a call fails, the operators raise.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb2q.so")

c_f32p = ctypes.POINTER(ctypes.c_float)
c_i32p = ctypes.POINTER(ctypes.c_int32)
_P = ctypes.c_void_p      # device / host data pointers are passed as integers
_I = ctypes.c_int
_L = ctypes.c_int64
_F = ctypes.c_float

REQ = {"null": 0, "write": 1, "inplace": 2, "add": 3}

UPD_STORE, UPD_EMA, UPD_GDRQ_WEIGHT, UPD_GDRQ_ACT, UPD_TWICE_STORE, UPD_TWICE_EMA = 1, 2, 3, 4, 5, 6
CLIP_NONE, CLIP_SYM, CLIP_WHERE_LE, CLIP_ZERO_T, CLIP_PACT, CLIP_WHERE_LT = 0, 1, 2, 3, 4, 5
MASK_OPEN, MASK_ABS_LE, MASK_LT = 1, 2, 3

# name -> argtypes (after the leading b2q_ctx*).  Every symbol of include/b2q.h is listed; tests check that
# the header, this table and the shared object agree.
_CTX_FUNCS = {
    "b2q_destroy": [],
    "b2q_num_sms": [],
    "b2q_stream_synchronize": [_P],
    "b2q_set_option": [ctypes.c_char_p, _I],
    "b2q_get_option": [ctypes.c_char_p, ctypes.POINTER(_I)],
    "b2q_timing_read": [_I, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                        ctypes.POINTER(_L), _I],
    "b2q_timing_read_range": [_I, ctypes.c_double, ctypes.c_double, ctypes.POINTER(ctypes.c_double),
                              ctypes.POINTER(ctypes.c_double), ctypes.POINTER(_L), _I],
    "b2q_absmax_f32": [_P, _L, _L, _L, _P, _P],
    "b2q_meanabs_f32": [_P, _L, _L, _L, _P, _P],
    "b2q_threshold_update_f32": [_I, _P, _P, _P, _L, _F, _F, _P],
    "b2q_qdq_f32": [_P, _P, _L, _L, _L, _P, _P, _F, _I, _I, _I, _P, _P, _P, _F, _P],
    "b2q_export_int8_f32": [_P, _P, _P, _L, _L, _L, _P, _F, _I, _P],
    "b2q_ste_bwd_f32": [_P, _P, _L, _I, _P],
    "b2q_zero_f32": [_P, _L, _P],
    "b2q_selftest": [_I, ctypes.POINTER(_L)],
    "b2q_peer_status": [_P, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)],
    "b2q_mask_bwd_f32": [_P, _P, _P, _L, _L, _L, _P, _F, _I, _I, _P],
    "b2q_minmax_quant_fwd_f32": [_I, _P, _P, _P, _L, _L, _I, _I, _I, _I, _F, _F, _I, _P],
    "b2q_minmax_quant_stat_f32": [_P, _L, _L, _I, _P, _P],
    "b2q_minmax_quant_finish_f32": [_I, _P, _P, _P, _P, _L, _L, _I, _I, _I, _I, _F, _F, _I, _P],
    "b2q_clipgrad_bwd_f32": [_P, _P, _P, _P, _L, _P],
    "b2q_gdrq_fwd_f32": [_P, _P, _P, _L, _L, _L, _I, _I, _I, _F, _F, _F, _I, _P],
    "b2q_gdrq_bwd_f32": [_P, _P, _P, _P, _L, _L, _L, _I, _P],
    "b2q_foldbn_data_fwd_f32": [_P, _P, _P, _L, _I, _F, _F, _P],
    "b2q_foldbn_weight_fwd_f32": [_P, _P, _P, _P, _P, _P, _P, _P, _F, _L, _L, _I, _I, _I, _P],
    "b2q_bn_batch_stats_f32": [_P, _L, _L, _L, _P, _P, _P],
    "b2q_bnstat_foldbn_weight_fwd_f32": [_P, _L, _L, _L, _P, _P, _P, _P, _P, _P, _P, _P, _F, _L, _I, _I, _I, _P],
    "b2q_clip_relu_fwd_f32": [_P, _P, _L, _F, _F, _I, _P],
    "b2q_wnq_fwd_f32": [_P, _P, _L, _L, _I, _F, _I, _P],
    "b2q_wnq_bwd_f32": [_P, _P, _P, _L, _L, _I, _I, _P],
    "b2q_pact_bwd_f32": [_P, _P, _P, _P, _P, _L, _I, _I, _I, _P],
    "b2q_dorefa_fwd_f32": [_P, _P, _P, _L, _F, _I, _P],
    "b2q_dorefa_bwd_f32": [_P, _P, _P, _P, _L, _I, _P],
    "b2q_qil_fwd_f32": [_I, _P, _P, _P, _P, _L, _F, _I, _P],
    "b2q_qil_bwd_f32": [_I, _P, _P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _P],
    "b2q_minmax_quant_fwd_host_f32": [_I, _P, _P, _P, _L, _L, _I, _I, _I, _I, _F, _F],
    "b2q_ste_bwd_host_f32": [_P, _P, _L],
    "b2q_clipgrad_bwd_host_f32": [_P, _P, _P, _P, _L],
    "b2q_host_sync": [],
    "b2q_peer_mailbox_create": [ctypes.POINTER(ctypes.c_void_p), _P],
    "b2q_peer_mailbox_open": [_P, ctypes.POINTER(ctypes.c_void_p)],
    "b2q_peer_mailbox_close": [_P],
    "b2q_peer_mailbox_destroy": [_P],
    "b2q_peer_buffer_create": [_L, ctypes.POINTER(ctypes.c_void_p), _P],
    "b2q_peer_allreduce_sum_f32": [ctypes.POINTER(ctypes.c_void_p), _L, _I, ctypes.POINTER(ctypes.c_void_p), _I, _I, _P],
    "b2q_peer_allreduce_max_f32": [ctypes.POINTER(ctypes.c_void_p), _L, ctypes.POINTER(ctypes.c_void_p), _I, _I, _P],
    "b2q_peer_minmax_quant_fwd_f32": [_I, _P, _P, _P, _L, _I, _F, _F, ctypes.POINTER(ctypes.c_void_p), _I, _I,
                                      ctypes.c_uint32, _P],
    "b2q_peer_meanabs_quant_fwd_f32": [_I, _P, _P, _P, _L, _F, _F, _F, ctypes.POINTER(ctypes.c_void_p), _I, _I, _P],
    "b2q_multi_plan_create": [_P, _I, ctypes.POINTER(ctypes.c_void_p)],
    "b2q_multi_plan_destroy": [_P],
    "b2q_multi_weight_quant_fwd_f32": [_P, _I, _I, _P],
    "b2q_multi_gdrq_weight_fwd_f32": [_P, _I, _I, _F, _F, _P],
    "b2q_multi_weight_ste_bwd_f32": [_P, _P],
}


class WeightDesc(ctypes.Structure):
    """b2q_weight_desc (include/b2q.h)."""
    _fields_ = [("x", ctypes.c_void_p), ("y", ctypes.c_void_p), ("aux", ctypes.c_void_p), ("dy", ctypes.c_void_p),
                ("dx", ctypes.c_void_p), ("rows", ctypes.c_int64), ("cols", ctypes.c_int64),
                ("per_channel", ctypes.c_int32), ("reserved", ctypes.c_int32)]
_PLAIN_FUNCS = {
    "b2q_abi_version": ([], _I),
    "b2q_peer_mailbox_bytes": ([], _I),
    "b2q_last_error": ([], ctypes.c_char_p),
    "b2q_create": ([_I, ctypes.POINTER(ctypes.c_void_p)], _I),
    "b2q_launch_count": ([ctypes.c_void_p], _L),
    # one process, N devices (b2q_comm): not bound to a single context
    "b2q_comm_create": ([ctypes.POINTER(ctypes.c_void_p), _I, ctypes.POINTER(ctypes.c_void_p)], _I),
    "b2q_comm_destroy": ([ctypes.c_void_p], _I),
    "b2q_comm_size": ([ctypes.c_void_p], _I),
    "b2q_comm_mailboxes": ([ctypes.c_void_p, _I, ctypes.POINTER(ctypes.POINTER(ctypes.c_void_p))], _I),
    "b2q_comm_allreduce_max_f32": ([ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), _L, ctypes.POINTER(ctypes.c_void_p)], _I),
    "b2q_comm_allreduce_sum_f32": ([ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), _L, _I, ctypes.POINTER(ctypes.c_void_p)], _I),
}
ALL_SYMBOLS = sorted(list(_CTX_FUNCS) + list(_PLAIN_FUNCS))


class B2QError(RuntimeError):
    pass


_lib = None
_lib_lock = threading.Lock()


def load():
    """Load libb2q.so (once).  Raises with build instructions when it is missing -- there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise B2QError(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
                "`make -C resnet.mxnet_b200/csrc` (needs nvcc with sm_100a support).  The quantization "
                "operators have no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, args in _CTX_FUNCS.items():
            fn = getattr(lib, name)
            fn.argtypes = [ctypes.c_void_p] + list(args)
            fn.restype = _I
        for name, (args, res) in _PLAIN_FUNCS.items():
            fn = getattr(lib, name)
            fn.argtypes = list(args)
            fn.restype = res
        _lib = lib
    return _lib


def last_error():
    msg = load().b2q_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


class Context(object):
    """Per-device handle (b2q_ctx).  Cheap to keep; holds only reduction workspaces."""

    def __init__(self, device):
        self.lib = load()
        self.device = int(device)
        h = ctypes.c_void_p()
        rc = self.lib.b2q_create(self.device, ctypes.byref(h))
        if rc != 0:
            raise B2QError("b2q_create(device=%d) failed: %s" % (self.device, last_error()))
        self.handle = h
        self._fn = {name: getattr(self.lib, name) for name in _CTX_FUNCS}

    def call(self, name, *args):
        rc = self._fn[name](self.handle, *args)
        if rc != 0:
            raise B2QError("%s failed (%d): %s" % (name, rc, last_error()))

    def launch_count(self):
        return int(self.lib.b2q_launch_count(self.handle))

    def num_sms(self):
        return int(self.lib.b2q_num_sms(self.handle))

    def set_option(self, key, value):
        self.call("b2q_set_option", key.encode(), int(value))

    def get_option(self, key):
        v = _I(0)
        self.call("b2q_get_option", key.encode(), ctypes.byref(v))
        return v.value

    def timing_read(self, kind=0, reset=False, min_bytes=0.0, max_bytes=0.0):
        """(device ms, algorithmic bytes, launches) recorded since the last reset for one kernel kind, optionally only
        the launches whose algorithmic bytes lie in [min_bytes, max_bytes)."""
        ms, by, n = ctypes.c_double(0), ctypes.c_double(0), _L(0)
        self.call("b2q_timing_read_range", int(kind), float(min_bytes), float(max_bytes), ctypes.byref(ms),
                  ctypes.byref(by), ctypes.byref(n), int(reset))
        return ms.value, by.value, n.value

    def host_sync(self):
        self.call("b2q_host_sync")

    def close(self):
        if self.handle:
            self.lib.b2q_destroy(self.handle)
            self.handle = None


_contexts = {}
_ctx_lock = threading.Lock()


def context(device):
    """Cached Context for a CUDA device index."""
    device = int(device)
    ctx = _contexts.get(device)
    if ctx is None:
        with _ctx_lock:
            ctx = _contexts.get(device)
            if ctx is None:
                ctx = Context(device)
                _contexts[device] = ctx
    return ctx


def total_launches():
    return sum(c.launch_count() for c in _contexts.values())
