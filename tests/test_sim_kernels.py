"""Kernel LOGIC parity on the GPU-less box: the CUDA sources compiled against tests/sim/cusim.h
(a cooperative-fiber model of a CUDA grid) versus the oracle.  This checks indexing, barriers,
warp collectives, table handling and arithmetic order -- not the hardware; the same cases run on
the real library in test_gpu_parity.py."""
import pytest

import parity_cases
from backends import SimBackend


@pytest.fixture(scope="module")
def be():
    return SimBackend()


@pytest.mark.parametrize("case", parity_cases.ALL_CASES, ids=lambda c: c.__name__)
def test_sim(case, be):
    case(be)
