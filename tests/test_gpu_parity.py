"""Parity proper: the sm_100a library, called through the C ABI, versus the oracle."""
import pytest

import parity_cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be():
    from backends import GpuBackend
    return GpuBackend()


@pytest.mark.parametrize("case", parity_cases.ALL_CASES, ids=lambda c: c.__name__)
def test_gpu(case, be):
    case(be)
