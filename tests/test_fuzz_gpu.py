"""Randomised parity sweep (tools/fuzz_parity.py) at a size that fits the suite: random k ranges, strand operations,
lengths, invalid bases, low-complexity rows, frozen subsets, binarize, both alphabets -- extraction bit exact,
gradient and loss within tolerance, against the oracle.  Needs a B200."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def test_randomised_parity_sweep(K, oracle):
    import fuzz_parity
    assert fuzz_parity.run(150, 3) == 0
