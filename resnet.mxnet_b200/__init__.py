"""b200quant -- B200-native (sm_100a) replacements for the int8 fake-quantization CustomOps of
XiaotaoChen/resnet.mxnet.  The directory is named ``resnet.mxnet_b200`` (not an importable identifier); it is
imported as ``b200quant`` through the loader module at the repository root.

Layout
    quant_ops, clip_grad_quantization_int8, int8_api, fold_bn_v1_gdrq, core_operator/   drop-ins for the
        reference modules of the same names (same classes, op_type strings, attributes, protocol)
    operator      the MXNet CustomOp protocol (real MXNet when importable, identical local one otherwise)
    _kernels      array-level calls into libb2q.so   |  _lib  ctypes binding  |  dlpack  zero-copy views
    harness       torch autograd driver for the operators (the only runnable host in this image)
    dist          cross-rank threshold synchronisation (allreduce-max) for data-parallel training
    csrc/         CUDA sources of libb2q.so (C ABI: include/b2q.h)
"""
from . import _lib  # noqa: F401
from .operator import REGISTRY, CustomOp, CustomOpProp, get_prop, register  # noqa: F401
from . import ops  # noqa: F401

__all__ = ["REGISTRY", "CustomOp", "CustomOpProp", "get_prop", "register", "ops"]
