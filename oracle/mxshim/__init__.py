"""TEST INFRASTRUCTURE ONLY -- a minimal stand-in for the ``mxnet`` Python package.

Purpose: let the reference's own, unmodified op classes (``/root/reference/symbol/quant_ops.py`` ...
``/root/reference/core/operator/*.py``) execute in a container where libmxnet cannot be installed, so that
their results can be recorded as golden vectors (``tests/golden/generate.py``) which then pin
``oracle/quant_oracle.py``.  It implements exactly the API surface those files touch and nothing more:

    mx.operator.{CustomOp, CustomOpProp, register}      mx.autograd.record
    mx.nd.{abs,max,mean,sum,round,clip,sign,where,reshape,swapaxes,sqrt,tanh,exp,power,zeros_like,
           ones_like,Convolution,array}                  mx.nd.NDArray (operators, indexing, asnumpy,
                                                         reshape, broadcast_like, attach_grad/grad/backward)
    mx.sym (oracle/mxshim/sym.py): a symbolic graph recorder for the reference's graph builders

Arrays are torch CPU float32 tensors (torch supplies IEEE float32 elementwise arithmetic and the autograd
the PACT/DoReFa/QIL ops replay).  The MXNet numerics it encodes are the [upstream] assumptions listed in
``oracle/quant_oracle.py`` (``round`` = roundf half-away, scalar operands applied in float32, ``clip`` by
comparisons, ``mean`` = rounded-once float32 sum / float32(N), full reductions return shape ``(1,)``).
It is NOT MXNet; fixtures made with it pin control flow, state machines, operator order and formulae of the
reference, under those stated numerics.

Usage:  ``import oracle.mxshim as shim; shim.install()``  then import reference files by path.
"""
import contextlib
import sys
import types

import numpy as np
import torch

torch.set_grad_enabled(False)  # like MXNet: nothing is recorded outside autograd.record()

_F32 = torch.float32


def _t(v, like=None):
    if isinstance(v, NDArray):
        return v._t
    if isinstance(v, torch.Tensor):
        return v
    if isinstance(v, np.ndarray):
        return torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
    return torch.tensor(float(v), dtype=_F32)  # python scalar -> float32 scalar operand


class NDArray(object):
    __array_priority__ = 1000.0

    def __init__(self, t):
        self._t = t

    # -- basic protocol ---------------------------------------------------------------------
    @property
    def shape(self):
        return tuple(self._t.shape)

    @property
    def size(self):
        return int(self._t.numel())

    @property
    def dtype(self):
        return np.float32

    def asnumpy(self):
        return self._t.detach().cpu().numpy().copy()

    def asscalar(self):
        return float(self._t.reshape(-1)[0])

    def copy(self):
        return NDArray(self._t.detach().clone())

    def __len__(self):
        return self._t.shape[0]

    def __repr__(self):
        return "<shim NDArray %s>\n%s" % (self.shape, self._t)

    # -- indexing: int index on a 1-d array yields shape (1,) like MXNet 1.x -----------------
    def __getitem__(self, idx):
        if isinstance(idx, int) and self._t.dim() == 1:
            return NDArray(self._t[idx:idx + 1])
        if isinstance(idx, NDArray):  # boolean-mask style used only as `q[mask] = 0`
            raise NotImplementedError
        return NDArray(self._t[idx])

    def __setitem__(self, idx, value):
        with torch.no_grad():
            if isinstance(idx, NDArray):
                self._t[idx._t != 0] = _t(value)
            elif isinstance(idx, int) and self._t.dim() == 1:
                self._t[idx:idx + 1] = _t(value)
            else:
                v = _t(value)
                tgt = self._t[idx]
                if v.dim() > 0 and v.numel() == tgt.numel():
                    v = v.reshape(tgt.shape)
                self._t[idx] = v

    # -- arithmetic (each operator = one float32 kernel) --------------------------------------
    def __neg__(self):
        return NDArray(-self._t)

    def __add__(self, o):
        return NDArray(self._t + _t(o))

    __radd__ = __add__

    def __iadd__(self, o):
        with torch.no_grad():
            self._t += _t(o)
        return self

    def __sub__(self, o):
        return NDArray(self._t - _t(o))

    def __rsub__(self, o):
        return NDArray(_t(o) - self._t)

    def __mul__(self, o):
        return NDArray(self._t * _t(o))

    __rmul__ = __mul__

    def __truediv__(self, o):
        return NDArray(self._t / _t(o))

    def __rtruediv__(self, o):
        return NDArray(_t(o) / self._t)

    __div__ = __truediv__
    __rdiv__ = __rtruediv__

    def __pow__(self, o):
        return NDArray(torch.pow(self._t, _t(o)))

    def _cmp(self, o, fn):
        return NDArray(fn(self._t, _t(o)).to(_F32))

    def __gt__(self, o):
        return self._cmp(o, torch.gt)

    def __ge__(self, o):
        return self._cmp(o, torch.ge)

    def __lt__(self, o):
        return self._cmp(o, torch.lt)

    def __le__(self, o):
        return self._cmp(o, torch.le)

    def __eq__(self, o):
        return self._cmp(o, torch.eq)

    def __ne__(self, o):
        return self._cmp(o, torch.ne)

    __hash__ = None

    # -- shape ops ---------------------------------------------------------------------------
    def reshape(self, *shape, **kw):
        if "shape" in kw:
            shape = kw["shape"]
        elif len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = shape[0]
        return NDArray(self._t.reshape(tuple(shape)))

    def broadcast_like(self, other):
        return NDArray(self._t.expand_as(_t(other)))

    @property
    def ndim(self):
        return self._t.dim()

    def expand_dims(self, axis):
        return NDArray(self._t.unsqueeze(axis))

    def __isub__(self, o):
        with torch.no_grad():
            self._t -= _t(o)
        return self

    def __itruediv__(self, o):
        with torch.no_grad():
            self._t /= _t(o)
        return self

    __idiv__ = __itruediv__

    # -- autograd ----------------------------------------------------------------------------
    def attach_grad(self):
        self._t = self._t.detach()
        self._t.grad = None
        self._t.requires_grad_(True)

    @property
    def grad(self):
        g = self._t.grad
        return NDArray(g if g is not None else torch.zeros_like(self._t))

    def backward(self, out_grad=None):
        with torch.enable_grad():
            self._t.backward(_t(out_grad) if out_grad is not None else torch.ones_like(self._t))


# ------------------------------------------------------------------------------------------------
# mx.nd functions
# ------------------------------------------------------------------------------------------------
def _axes(axis, ndim):
    if axis is None or axis == ():
        return None
    if isinstance(axis, int):
        return (axis,)
    return tuple(axis)


def _reduce_shape1(t):
    return t.reshape(1) if t.dim() == 0 else t


def nd_abs(x):
    return NDArray(torch.abs(_t(x)))


def nd_sign(x):
    return NDArray(torch.sign(_t(x)))


def nd_sqrt(x):
    return NDArray(torch.sqrt(_t(x)))


def nd_tanh(x):
    return NDArray(torch.tanh(_t(x)))


def nd_exp(x):
    return NDArray(torch.exp(_t(x)))


def nd_power(x, p):
    return NDArray(torch.pow(_t(x), _t(p)))


def nd_max(x, axis=None):
    t = _t(x)
    ax = _axes(axis, t.dim())
    return NDArray(_reduce_shape1(torch.amax(t, dim=ax) if ax is not None else torch.max(t)))


def nd_sum(x, axis=None):
    t = _t(x)
    ax = _axes(axis, t.dim())
    if t.requires_grad and torch.is_grad_enabled():
        s = torch.sum(t, dim=ax) if ax is not None else torch.sum(t)
    else:  # rounded-once float32 sum (model of libmxnet's Kahan-compensated reduce)
        s = (torch.sum(t.double(), dim=ax) if ax is not None else torch.sum(t.double())).to(_F32)
    return NDArray(_reduce_shape1(s))


def nd_mean(x, axis=None):
    t = _t(x)
    s = nd_sum(x, axis=axis)._t
    n = t.numel() // max(s.numel(), 1)
    return NDArray(s / torch.tensor(float(n), dtype=_F32))


def nd_round(x):
    """roundf (half away from zero)."""
    t = _t(x)
    tr = torch.trunc(t)
    r = t - tr
    out = tr + torch.where(torch.abs(r) >= 0.5, torch.sign(t), torch.zeros_like(t))
    return NDArray(torch.copysign(out, t))


def nd_clip(x, a_min, a_max):
    t = _t(x)
    lo = torch.tensor(float(a_min), dtype=_F32)
    hi = torch.tensor(float(a_max), dtype=_F32)
    return NDArray(torch.where(t > hi, hi, torch.where(t < lo, lo, t)))


def nd_where(cond, x, y):
    return NDArray(torch.where(_t(cond) != 0, _t(x), _t(y)))


def nd_reshape(x, shape=None):
    return NDArray(_t(x).reshape(tuple(shape)))


def nd_swapaxes(x, dim1=0, dim2=1):
    return NDArray(_t(x).transpose(dim1, dim2).contiguous())


def nd_zeros_like(x):
    return NDArray(torch.zeros_like(_t(x)))


def nd_ones_like(x):
    return NDArray(torch.ones_like(_t(x)))


def nd_array(a, ctx=None, dtype=None):
    return NDArray(torch.from_numpy(np.array(a, dtype=np.float32)))


def nd_convolution(data=None, weight=None, bias=None, kernel=None, stride=(1, 1), pad=(0, 0), dilate=(1, 1),
                   num_filter=None, num_group=1, no_bias=False, name=None, **_):
    y = torch.nn.functional.conv2d(_t(data).double(), _t(weight).double(), None, stride=tuple(stride),
                                   padding=tuple(pad), dilation=tuple(dilate), groups=int(num_group)).to(_F32)
    if bias is not None and not no_bias:
        y = y + _t(bias).reshape(1, -1, 1, 1)
    return NDArray(y)


# ------------------------------------------------------------------------------------------------
# mx.operator
# ------------------------------------------------------------------------------------------------
class CustomOp(object):
    def forward(self, is_train, req, in_data, out_data, aux):
        pass

    def backward(self, req, out_grad, in_data, out_data, in_grad, aux):
        pass

    def assign(self, dst, req, src):
        if req == "null":
            return
        elif req in ("write", "inplace"):
            dst[:] = src
        elif req == "add":
            dst[:] += src


class CustomOpProp(object):
    def __init__(self, need_top_grad=True):
        self.need_top_grad_ = need_top_grad


REGISTRY = {}


def register(reg_name):
    def deco(prop_cls):
        REGISTRY[reg_name] = prop_cls
        return prop_cls
    return deco


@contextlib.contextmanager
def _record(train_mode=True):
    with torch.enable_grad():
        yield


def install():
    """Put the shim in sys.modules as ``mxnet`` (refuses to shadow a real installation)."""
    if "mxnet" in sys.modules and not getattr(sys.modules["mxnet"], "__is_b2q_shim__", False):
        raise RuntimeError("a real mxnet is already imported; the shim must not shadow it")
    mx = types.ModuleType("mxnet")
    mx.__is_b2q_shim__ = True
    nd = types.ModuleType("mxnet.nd")
    for name, fn in dict(abs=nd_abs, sign=nd_sign, sqrt=nd_sqrt, tanh=nd_tanh, exp=nd_exp, power=nd_power,
                         max=nd_max, sum=nd_sum, mean=nd_mean, round=nd_round, clip=nd_clip, where=nd_where,
                         reshape=nd_reshape, swapaxes=nd_swapaxes, zeros_like=nd_zeros_like,
                         ones_like=nd_ones_like, array=nd_array, Convolution=nd_convolution,
                         NDArray=NDArray).items():
        setattr(nd, name, fn)
    op = types.ModuleType("mxnet.operator")
    op.CustomOp, op.CustomOpProp, op.register = CustomOp, CustomOpProp, register
    ag = types.ModuleType("mxnet.autograd")
    ag.record = _record
    init = types.ModuleType("mxnet.init")

    class Initializer(object):
        """[upstream python/mxnet/initializer.py]: ``dumps()`` is what ends up in a variable's ``__init__`` attribute."""

        def __init__(self, **kwargs):
            self._kwargs = kwargs

        def dumps(self):
            import json
            return json.dumps([self.__class__.__name__.lower(), self._kwargs])

    class Constant(Initializer):
        def __init__(self, value):
            super(Constant, self).__init__(value=value)

    init.Initializer, init.Constant = Initializer, Constant
    from . import sym as _sym
    _sym.set_registries(REGISTRY)
    mx.nd, mx.ndarray, mx.operator, mx.autograd, mx.init = nd, nd, op, ag, init
    mx.sym = mx.symbol = _sym
    sys.modules.update({"mxnet": mx, "mxnet.nd": nd, "mxnet.operator": op, "mxnet.autograd": ag,
                        "mxnet.init": init, "mxnet.sym": _sym, "mxnet.symbol": _sym})
    return mx
