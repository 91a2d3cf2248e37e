"""Randomised parity run (-m gpu): tools/fuzz_parity.py with a fixed seed -- random sizes biased towards tile and chunk
boundaries, levels 0..10, every quantizer, both interpolators, batches, noise / photograph-like / saturated / (x*y)&255
content, packed planes through the host API and padded rows through the device API -- against the C oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("seed", [101, 202])
def test_fuzz_parity(seed):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_parity.py"), "--cases", "120", "--seed", str(seed)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "fuzz ok: 120 cases" in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]
