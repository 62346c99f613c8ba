"""Drop-ins for the reference's ``core/operator`` package (GDRQ, PACT, WNQ, QIL, QIL_V2, QIL_V3, quant_utils)."""
from . import GDRQ, PACT, QIL, QIL_V2, QIL_V3, WNQ, quant_utils  # noqa: F401
