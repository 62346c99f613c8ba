"""CPU-side checks of the boundary: the shared library loads, exports exactly the symbols include/b2q.h
declares, the ctypes table agrees, and the product path refuses to run without a GPU (no CPU fallback)."""
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b2q.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2q_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    from b200quant import _lib
    assert os.path.exists(_lib.LIB_PATH), "build libb2q.so first (python -c 'import __graft_entry__ as g; g.build()')"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r" T (b2q_[a-z0-9_]+)", out)))
    assert header_symbols() == exported


def test_ctypes_table_matches_header():
    from b200quant import _lib
    assert header_symbols() == _lib.ALL_SYMBOLS
    lib = _lib.load()
    assert lib.b2q_abi_version() == 1
    for name in _lib.ALL_SYMBOLS:
        assert hasattr(lib, name)


def test_header_cites_reference_for_every_entry_point():
    src = open(HEADER).read()
    for ref_file in ("symbol/quant_ops.py", "symbol/clip_grad_quantization_int8.py", "symbol/fold_bn_v1_gdrq.py",
                     "core/operator/GDRQ.py", "core/operator/PACT.py", "core/operator/WNQ.py", "core/operator/QIL.py"):
        assert ref_file in src


def test_no_cpu_fallback():
    """without a CUDA device the operators must raise, not compute on the CPU"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import b200quant
    from b200quant._lib import B2QError
    op = b200quant.get_prop("Quantization_int8_V2")(quant_mode="minmax", is_weight="True").create_operator(None, None, None)
    x = torch.randn(8, 3, 3, 3)
    with pytest.raises(B2QError):
        op.forward(True, ["write"], [x], [torch.empty_like(x)], [torch.ones(1)])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "resnet.mxnet_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "quant_oracle" not in text, f


def test_dlpack_view_is_zero_copy():
    import torch
    from b200quant.dlpack import as_buffer, _from_capsule
    a = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    b = as_buffer(a)
    assert b.ptr == a.ctypes.data and b.shape == (2, 3, 4) and b.numel == 24 and not b.on_device
    t = torch.arange(10, dtype=torch.float32)[2:]
    c = _from_capsule(t.__dlpack__())
    assert c.ptr == t.data_ptr() and c.shape == (8,)
    assert as_buffer(t).ptr == t.data_ptr()
    with pytest.raises(TypeError):
        as_buffer(np.zeros(3, dtype=np.float64))
    with pytest.raises(ValueError):
        as_buffer(torch.zeros(4, 4).t())
    with pytest.raises(ValueError):
        as_buffer(np.zeros((4, 4), dtype=np.float32).T)
