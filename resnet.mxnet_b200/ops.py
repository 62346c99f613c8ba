"""Importing this module registers every operator of the hot path (like importing the reference's
``symbol`` package and ``core/graph_optimize.py`` does, graph_optimize.py:22-27)."""
from . import clip_grad_quantization_int8, fold_bn_v1_gdrq, quant_ops  # noqa: F401
from .core_operator import GDRQ, PACT, QIL, QIL_V2, QIL_V3, WNQ  # noqa: F401
from .operator import REGISTRY  # noqa: F401

OP_TYPES = ("Quantization_int8_V2", "ClipGrad_Quantization_int8", "GDRQ_Fold_BN", "GDRQ_PY", "CLIP_RELU_PY",
            "QUANT_STE_PY", "PACT_PY", "PACT_V2_PY", "DoReFa_PY", "WNQ_PY", "QIL_PY", "QIL_V2_PY", "QIL_V3_PY")
