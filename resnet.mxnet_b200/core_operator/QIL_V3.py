"""Drop-in for ``core/operator/QIL_V3.py``: op_type ``QIL_V3_PY`` (pp = exp(ep), distance = exp(ed))."""
from ..operator import register
from .QIL import QIL_PY, _QILProp


class QIL_V3_PY(QIL_PY):
    VARIANT = 3


@register("QIL_V3_PY")
class QIL_V3_PYProp(_QILProp):
    OP = QIL_V3_PY
    ARGS = ["data", "ep", "ed", "gamma"]
