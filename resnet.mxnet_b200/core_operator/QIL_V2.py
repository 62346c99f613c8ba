"""Drop-in for ``core/operator/QIL_V2.py``: op_type ``QIL_V2_PY`` (centre / distance parametrisation)."""
from ..operator import register
from .QIL import QIL_PY, _QILProp


class QIL_V2_PY(QIL_PY):
    VARIANT = 2


@register("QIL_V2_PY")
class QIL_V2_PYProp(_QILProp):
    OP = QIL_V2_PY
    ARGS = ["data", "center", "distance", "gamma"]
