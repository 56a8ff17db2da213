"""A bounded slice of the fuzz campaign (tests/studies/fuzz_sim.py): random small inputs through the kernel sources on
the CPU grid simulator, exact comparisons with the oracle.  Fixed seeds, a fixed number of iterations per target."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "studies"))
import fuzz_sim  # noqa: E402


@pytest.fixture(scope="module")
def be():
    from backends import SimBackend
    return SimBackend()


@pytest.mark.parametrize("target", fuzz_sim.TARGETS, ids=lambda f: f.__name__)
def test_fuzz_slice(target, be):
    for it in range(25):
        target(be, np.random.default_rng([2025, it]))
