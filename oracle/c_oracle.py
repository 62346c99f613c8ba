"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/c/libb2q_oracle.so (the C/OpenMP restatement of the
reference's CPU execution: one pass and one temporary per mx.nd call).  Used by tests (checked against
quant_oracle.py) and by bench.py's cpu_baseline / --impl reference legs as the timed reference CPU path."""
import ctypes
import os
import subprocess

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c")
_PATH = os.path.join(_DIR, "libb2q_oracle.so")
_lib = None
_P, _I, _L, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
REQ = {"null": 0, "write": 1, "inplace": 2, "add": 3}


def load(build_if_missing=True):
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_PATH) and build_if_missing:
        subprocess.run(["make", "-C", _DIR], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    lib = ctypes.CDLL(_PATH)
    lib.b2qo_num_threads.restype = _I
    lib.b2qo_set_num_threads.argtypes = [_I]
    lib.b2qo_minmax_quant_fwd.argtypes = [_I, _P, _P, _P, _L, _L, _I, _I, _I, _I, _F, _F, _I]
    lib.b2qo_ste_bwd.argtypes = [_P, _P, _L, _I]
    lib.b2qo_clipgrad_bwd.argtypes = [_P, _P, _P, _P, _L]
    lib.b2qo_gdrq_fwd.argtypes = [_P, _P, _P, _L, _I, _I, _I, _F, _F, _F, _I]
    lib.b2qo_gdrq_bwd.argtypes = [_P, _P, _P, _P, _L, _I]
    lib.b2qo_foldbn_data_fwd.argtypes = [_P, _P, _P, _L, _I, _F, _F]
    lib.b2qo_foldbn_weight_fwd.argtypes = [_P, _P, _P, _P, _P, _P, _P, _P, _F, _L, _L, _I, _I, _I]
    _lib = lib
    return lib


def _p(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def _rc(shape):
    rows = int(shape[0])
    return rows, int(np.prod(shape)) // rows


def num_threads():
    return int(load().b2qo_num_threads())


def set_num_threads(n):
    """Thread count of the OpenMP loops (launchers such as torchrun export OMP_NUM_THREADS=1 to their workers)."""
    load().b2qo_set_num_threads(int(n))
    return num_threads()


def use_all_host_threads():
    import os
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        n = os.cpu_count() or 1
    return set_num_threads(n)


def minmax_quant_fwd(variant, x, y, aux, is_weight, per_channel, is_train, init, ema_decay, req="write"):
    rows, cols = _rc(x.shape)
    rc = load().b2qo_minmax_quant_fwd(variant, _p(x), _p(y), _p(aux), rows, cols, int(is_weight), int(per_channel),
                                      int(is_train), int(init), float(np.float32(ema_decay)),
                                      float(np.float32(1 - ema_decay)), REQ[req])
    assert rc == 0


def ste_bwd(dy, dx, req="write"):
    assert load().b2qo_ste_bwd(_p(dy), _p(dx), dy.size, REQ[req]) == 0


def clipgrad_bwd(x, dy, dx, aux):
    assert load().b2qo_clipgrad_bwd(_p(x), _p(dy), _p(dx), _p(aux), x.size) == 0


def gdrq_fwd(x, y, alpha, is_weight, fix_alpha, do_round, qlevel, ktimes, lamda, req="write"):
    assert load().b2qo_gdrq_fwd(_p(x), _p(y), _p(alpha), x.size, int(is_weight), int(fix_alpha), int(do_round),
                                float(np.float32(qlevel)), float(np.float32(ktimes)), float(np.float32(lamda)),
                                REQ[req]) == 0


def gdrq_bwd(x, dy, dx, alpha, req="write"):
    assert load().b2qo_gdrq_bwd(_p(x), _p(dy), _p(dx), _p(alpha), x.size, REQ[req]) == 0


def foldbn_data_fwd(x, y, aux, init, ema_decay):
    assert load().b2qo_foldbn_data_fwd(_p(x), _p(y), _p(aux), x.size, int(init), float(np.float32(ema_decay)),
                                       float(np.float32(1 - ema_decay))) == 0


def foldbn_weight_fwd(w, wq, bias, aux, gamma, beta, mean, var, eps, per_channel, quantize, is_train):
    cout, cols = _rc(w.shape)
    assert load().b2qo_foldbn_weight_fwd(_p(w), _p(wq), _p(bias), _p(aux), _p(gamma), _p(beta), _p(mean), _p(var),
                                         float(np.float32(eps)), cout, cols, int(per_channel), int(quantize),
                                         int(is_train)) == 0
