"""Checkpoint interchange with the reference.

The reference saves ``prefix-%04d.params`` with ``mx.callback.do_checkpoint`` (train.py:218, core/solver.py:173-175)
and resumes with ``mx.model.load_checkpoint`` (train.py:224-227): one MXNet NDArray-list file whose entries are named
``arg:<param>`` / ``aux:<state>`` -- so the quantization thresholds (``*_minmax``, ``*_alpha``, ``*_gamma``,
``*_pruning_point``, ``*_clipping_point``) travel with the weights.  This module reads and writes that container for
dense float32 / float64 / float16 / int32 / int64 / uint8 / int8 CPU arrays, following the layout of MXNet 1.x's
``NDArray::Save`` / ``NDArray::Load`` (src/ndarray/ndarray.cc [upstream]; libmxnet is not installable in this image, so
the layout is restated from the source and round-trip tested, not verified against a real binary):

    uint64 0x112, uint64 0, uint64 count, count x NDArray, uint64 count, count x (uint64 len, bytes)
    NDArray (V2) = uint32 0xF993fac9, int32 stype(0), uint32 ndim, int64 dims[ndim], int32 dev_type(1=cpu),
                   int32 dev_id, int32 type_flag, raw data            (V1 = 0xF993fac8 has no stype, uint32 dims)

What the reference forgets to checkpoint -- each operator's ``delay_quant`` countdown and first-batch ``init`` flag
 -- goes into a JSON side-car ``<file>.opstate.json``.
"""
import json
import struct

import numpy as np

LIST_MAGIC = 0x112
V1_MAGIC, V2_MAGIC, V3_MAGIC = 0xF993FAC8, 0xF993FAC9, 0xF993FACA
_TYPES = {0: np.float32, 1: np.float64, 2: np.float16, 3: np.uint8, 4: np.int32, 5: np.int8, 6: np.int64}
_FLAGS = {np.dtype(v): k for k, v in _TYPES.items()}


def _to_numpy(a):
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    elif hasattr(a, "asnumpy"):
        a = a.asnumpy()
    return np.ascontiguousarray(a)


def _write_ndarray(f, a):
    a = _to_numpy(a)
    if a.dtype not in _FLAGS:
        raise TypeError("unsupported dtype %s" % a.dtype)
    f.write(struct.pack("<Ii", V2_MAGIC, 0))
    f.write(struct.pack("<I", a.ndim))
    f.write(struct.pack("<%dq" % a.ndim, *a.shape))
    f.write(struct.pack("<iii", 1, 0, _FLAGS[a.dtype]))
    f.write(a.tobytes())


def _read_ndarray(f):
    (magic,) = struct.unpack("<I", f.read(4))
    if magic in (V2_MAGIC, V3_MAGIC):
        (stype,) = struct.unpack("<i", f.read(4))
        if stype != 0:
            raise NotImplementedError("sparse NDArrays are not supported")
        (ndim,) = struct.unpack("<I" if magic == V2_MAGIC else "<i", f.read(4))
        shape = struct.unpack("<%dq" % ndim, f.read(8 * ndim)) if ndim > 0 else ()
    elif magic == V1_MAGIC:
        (ndim,) = struct.unpack("<I", f.read(4))
        shape = struct.unpack("<%dI" % ndim, f.read(4 * ndim)) if ndim > 0 else ()
    else:   # legacy: the word just read is ndim of a uint32 shape
        ndim = magic
        shape = struct.unpack("<%dI" % ndim, f.read(4 * ndim)) if ndim > 0 else ()
    if ndim == 0:
        return np.zeros((), np.float32)
    f.read(8)   # context (dev_type, dev_id)
    (flag,) = struct.unpack("<i", f.read(4))
    dt = np.dtype(_TYPES[flag])
    n = int(np.prod(shape))
    return np.frombuffer(f.read(n * dt.itemsize), dtype=dt).reshape(shape).copy()


def save_ndarray_dict(path, named):
    """Write ``{name: array}`` as an MXNet NDArray-list file (mx.nd.save)."""
    names = list(named.keys())
    with open(path, "wb") as f:
        f.write(struct.pack("<QQ", LIST_MAGIC, 0))
        f.write(struct.pack("<Q", len(names)))
        for k in names:
            _write_ndarray(f, named[k])
        f.write(struct.pack("<Q", len(names)))
        for k in names:
            b = k.encode("utf-8")
            f.write(struct.pack("<Q", len(b)))
            f.write(b)


def load_ndarray_dict(path):
    """Read an MXNet NDArray-list file (mx.nd.load) into ``{name: numpy array}``."""
    with open(path, "rb") as f:
        magic, _ = struct.unpack("<QQ", f.read(16))
        if magic != LIST_MAGIC:
            raise ValueError("%s is not an MXNet NDArray list (magic %#x)" % (path, magic))
        (count,) = struct.unpack("<Q", f.read(8))
        arrays = [_read_ndarray(f) for _ in range(count)]
        (ncount,) = struct.unpack("<Q", f.read(8))
        names = []
        for _ in range(ncount):
            (ln,) = struct.unpack("<Q", f.read(8))
            names.append(f.read(ln).decode("utf-8"))
    if ncount == 0:
        names = [str(i) for i in range(count)]
    return dict(zip(names, arrays))


def save_checkpoint(prefix, epoch, arg_params, aux_params, op_state=None):
    """``prefix-%04d.params`` as mx.model.save_checkpoint writes it (+ the op-state side-car)."""
    path = "%s-%04d.params" % (prefix, epoch)
    named = {"arg:" + k: v for k, v in arg_params.items()}
    named.update({"aux:" + k: v for k, v in aux_params.items()})
    save_ndarray_dict(path, named)
    if op_state is not None:
        with open(path + ".opstate.json", "w") as f:
            json.dump(op_state, f, indent=1, sort_keys=True)
    return path


def load_checkpoint(prefix, epoch):
    """(arg_params, aux_params, op_state) from ``prefix-%04d.params`` (mx.model.load_checkpoint without the symbol)."""
    path = "%s-%04d.params" % (prefix, epoch)
    arg_params, aux_params = {}, {}
    for k, v in load_ndarray_dict(path).items():
        kind, _, name = k.partition(":")
        (arg_params if kind == "arg" else aux_params)[name] = v
    try:
        with open(path + ".opstate.json") as f:
            op_state = json.load(f)
    except IOError:
        op_state = None
    return arg_params, aux_params, op_state


def restore_quant_layers(model, arg_params, aux_params, op_state=None, allow_missing=True):
    """Load a checkpoint into a torch model built from ``harness.QuantConv2d / QuantLinear`` (the reference resumes
    with allow_missing=True so that an fp32 checkpoint can seed a quantized graph, config/edict_config.py:27)."""
    import torch
    from .harness import _QuantLayer
    missing = []
    for mod in model.modules():
        if not isinstance(mod, _QuantLayer):
            continue
        args, _ = mod.mx_names()
        for name, param in args.items():
            if name in arg_params:
                with torch.no_grad():
                    param.copy_(torch.from_numpy(np.asarray(arg_params[name])).to(param.device))
            else:
                missing.append(name)
        for node, q in ((mod.weight_node_name, mod.weight_quant), (mod.data_node_name, mod.data_quant)):
            for aname in q.aux_names:
                key = node + "_" + aname
                if key in aux_params:
                    t = torch.from_numpy(np.asarray(aux_params[key], dtype=np.float32))
                    dev = next(model.parameters()).device
                    if getattr(q, aname, None) is None:
                        q.register_buffer(aname, t.to(dev).clone())
                    else:
                        getattr(q, aname).copy_(t.to(getattr(q, aname).device))
                    q._aux_ready = all(getattr(q, n, None) is not None for n in q.aux_names)
                else:
                    missing.append(key)
            if op_state and node in op_state:
                q.set_extra_state(op_state[node])
    if missing and not allow_missing:
        raise KeyError("missing in checkpoint: %s" % ", ".join(missing))
    return missing
