"""GPU: a short run of the randomised parity generator (tools/fuzz_parity.py: random field sizes, cell sizes, origins up
to 1e4, triangle size mixes, grid-snapped and axis-aligned geometry, degenerate triangles, random agent parameters),
every output array bit for bit against the oracle.  Longer runs: python tools/fuzz_parity.py 300 <seed>."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [101, 202])
def test_random_scenes_match_oracle(seed, monkeypatch):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import fuzz_parity
    monkeypatch.setattr(sys, "argv", ["fuzz_parity.py", "60", str(seed)])
    assert fuzz_parity.main() == 0
    os.environ["RCB_FORCE_GENERIC"] = "0"
