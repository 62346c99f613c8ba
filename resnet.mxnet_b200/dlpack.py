"""Zero-copy views of framework arrays for the C ABI.

``as_buffer(obj)`` returns ``Buf(ptr, shape, numel, on_device, device_id, keep)`` for
  * a torch tensor (fast path: ``data_ptr()``),
  * an MXNet NDArray (``to_dlpack_for_read/write`` [upstream python/mxnet/ndarray/ndarray.py]),
  * anything else exporting ``__dlpack__`` (numpy arrays, cupy, ...), parsed straight from the capsule's
    ``DLManagedTensor`` with ctypes -- no copy, no framework dependency.
Only dense float32 row-major arrays are accepted, which is what the reference operators produce
(``infer_type`` passes float32 through, symbol/quant_ops.py:66-67).
"""
import collections
import ctypes

try:  # torch supplies device memory and streams in the harness; the operators work without it (DLPack path)
    import torch
except Exception:  # pragma: no cover
    torch = None

Buf = collections.namedtuple("Buf", "ptr shape numel on_device device_id keep")

kDLCPU, kDLCUDA, kDLCUDAHost, kDLCUDAManaged = 1, 2, 3, 13
kDLFloat = 2


class DLDevice(ctypes.Structure):
    _fields_ = [("device_type", ctypes.c_int32), ("device_id", ctypes.c_int32)]


class DLDataType(ctypes.Structure):
    _fields_ = [("code", ctypes.c_uint8), ("bits", ctypes.c_uint8), ("lanes", ctypes.c_uint16)]


class DLTensor(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("device", DLDevice), ("ndim", ctypes.c_int32),
                ("dtype", DLDataType), ("shape", ctypes.POINTER(ctypes.c_int64)),
                ("strides", ctypes.POINTER(ctypes.c_int64)), ("byte_offset", ctypes.c_uint64)]


class DLManagedTensor(ctypes.Structure):
    _fields_ = [("dl_tensor", DLTensor), ("manager_ctx", ctypes.c_void_p), ("deleter", ctypes.c_void_p)]


class DLPackVersion(ctypes.Structure):
    _fields_ = [("major", ctypes.c_uint32), ("minor", ctypes.c_uint32)]


class DLManagedTensorVersioned(ctypes.Structure):
    _fields_ = [("version", DLPackVersion), ("manager_ctx", ctypes.c_void_p), ("deleter", ctypes.c_void_p),
                ("flags", ctypes.c_uint64), ("dl_tensor", DLTensor)]


_api = ctypes.pythonapi
_api.PyCapsule_IsValid.argtypes = [ctypes.py_object, ctypes.c_char_p]
_api.PyCapsule_IsValid.restype = ctypes.c_int
_api.PyCapsule_GetPointer.argtypes = [ctypes.py_object, ctypes.c_char_p]
_api.PyCapsule_GetPointer.restype = ctypes.c_void_p


def _from_capsule(capsule):
    if _api.PyCapsule_IsValid(capsule, b"dltensor"):
        p = _api.PyCapsule_GetPointer(capsule, b"dltensor")
        t = ctypes.cast(p, ctypes.POINTER(DLManagedTensor)).contents.dl_tensor
    elif _api.PyCapsule_IsValid(capsule, b"dltensor_versioned"):
        p = _api.PyCapsule_GetPointer(capsule, b"dltensor_versioned")
        t = ctypes.cast(p, ctypes.POINTER(DLManagedTensorVersioned)).contents.dl_tensor
    else:
        raise TypeError("not an unconsumed DLPack capsule")
    if not (t.dtype.code == kDLFloat and t.dtype.bits == 32 and t.dtype.lanes == 1):
        raise TypeError("b200quant operators take float32 tensors (got dtype code=%d bits=%d)"
                        % (t.dtype.code, t.dtype.bits))
    shape = tuple(int(t.shape[i]) for i in range(t.ndim))
    if t.strides:  # NULL strides mean compact row-major
        expect = 1
        for i in range(t.ndim - 1, -1, -1):
            if shape[i] != 1 and int(t.strides[i]) != expect:
                raise ValueError("b200quant operators take dense row-major tensors (shape %s strides %s)"
                                 % (shape, tuple(int(t.strides[i]) for i in range(t.ndim))))
            expect *= shape[i]
    numel = 1
    for s in shape:
        numel *= s
    dev = t.device.device_type
    if dev not in (kDLCPU, kDLCUDA, kDLCUDAHost, kDLCUDAManaged):
        raise TypeError("unsupported DLPack device type %d" % dev)
    on_device = dev in (kDLCUDA, kDLCUDAManaged)
    ptr = (t.data or 0) + int(t.byte_offset)
    return Buf(ptr, shape, numel, on_device, int(t.device.device_id), capsule)


def as_buffer(obj, write=False):
    """Zero-copy view of ``obj``; raises for non-float32 / non-dense arrays (never copies silently)."""
    if torch is not None and isinstance(obj, torch.Tensor):
        if obj.dtype != torch.float32:
            raise TypeError("b200quant operators take float32 tensors (got %s)" % obj.dtype)
        if not obj.is_contiguous():
            raise ValueError("b200quant operators take dense row-major tensors; call .contiguous() first")
        return Buf(obj.data_ptr(), tuple(obj.shape), obj.numel(), obj.is_cuda,
                   obj.device.index if obj.is_cuda else 0, obj)
    if hasattr(obj, "to_dlpack_for_write"):  # mx.nd.NDArray: these calls wait for pending engine reads/writes
        return _from_capsule(obj.to_dlpack_for_write() if write else obj.to_dlpack_for_read())
    if hasattr(obj, "__dlpack__"):
        return _from_capsule(obj.__dlpack__())
    raise TypeError("cannot view %r as a DLPack tensor" % type(obj))


FOREIGN_DEVICES = set()   # devices whose legacy default stream got launches for arrays torch does not own


def current_stream(buf):
    """CUDA stream to launch on for ``buf``'s device: torch's current stream when torch owns the memory,
    the legacy default stream (0) otherwise (MXNet: ``operator.register`` wraps forward / backward so that they
    synchronise that stream before returning to the engine)."""
    if torch is not None and isinstance(buf.keep, torch.Tensor) and buf.on_device:
        return _raw_stream(buf.device_id)
    if buf.on_device:
        FOREIGN_DEVICES.add(buf.device_id)
    return 0


def sync_foreign():
    """Wait for the default-stream launches made for non-torch arrays (no-op when there were none)."""
    if not FOREIGN_DEVICES:
        return
    from . import _lib
    while FOREIGN_DEVICES:
        _lib.context(FOREIGN_DEVICES.pop()).call("b2q_stream_synchronize", 0)


def _raw_stream(device_id):
    try:
        return torch._C._cuda_getCurrentRawStream(device_id)
    except AttributeError:  # pragma: no cover - older torch
        return torch.cuda.current_stream(device_id).cuda_stream
