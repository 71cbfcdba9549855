"""Random layer configurations (mode, concat / mean, G, F, K, P, B, N, bias, GSO dtype) through path="auto": forward, output
strides, attention and every gradient against the CPU oracle (tools/fuzz_parity.py; fixed seeds).  Guards the dispatch: a
*_supported predicate that promises a layout its kernel cannot take (K = 4 once did) fails here."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [3, 11])
def test_random_configurations_match_the_oracle(seed):
    import fuzz_parity
    failures = fuzz_parity.run(60, seed, verbose=False)
    assert not failures, "\n".join(failures)
