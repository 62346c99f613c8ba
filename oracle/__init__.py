"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's fake-quant CustomOps.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import, call, link or execute it, and there
only as the checker (or as the timed CPU baseline), never as the thing shipped.  The product path
(``resnet.mxnet_b200`` a.k.a. ``b200quant``) never imports this package and raises when its CUDA
library is missing.

Contents
--------
quant_oracle.py   NumPy restatement of every op on the hot path, one class per
                  reference CustomOp with the same forward/backward protocol, citing reference file:line.
mxshim/           A tiny emulation of the ``mxnet`` Python API surface the reference op files use, backed
                  by torch-CPU fp32 tensors.  It lets the reference's *own, unmodified* op classes run in
                  this container (``/root/reference`` is imported, never copied) to generate the golden
                  vectors under ``tests/golden/`` that pin ``quant_oracle.py``.
c/                Plain-C (OpenMP) restatement with MXNet's kernel-per-expression structure; it is the
                  timed CPU baseline and is itself checked against quant_oracle.py.

Parity status: the true MXNet binary is not installable here, so the oracle is pinned
against the reference op sources executed over ``mxshim`` (whose numerics encode the [upstream]
assumptions listed in quant_oracle.py), not against a real libmxnet; and against the NumPy simulators the
reference ships next to its ops (tests/golden/simulators.npz, tests/test_simulators.py).  The fork's C++ contrib ops
(``contrib.Quantization_int8`` etc.) have no source in the reference tree: parity unpinned.
"""
