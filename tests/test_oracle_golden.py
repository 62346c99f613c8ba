"""The NumPy oracle must reproduce what the reference's own op classes produced (over oracle/mxshim)."""
import numpy as np
import pytest

from oracle import quant_oracle as qo
from tests.golden_util import CASES, MANIFEST, check_against_golden


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_oracle_matches_reference_fixture(case):
    op = qo.create(case["op_type"], **case["attrs"])
    check_against_golden(case, op, lambda a: np.array(a, dtype=np.float32), lambda a: np.array(a))


def test_manifest_covers_every_case():
    assert sorted(MANIFEST.keys()) == sorted(c["id"] for c in CASES)
