import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def load_plane(name):
    from PIL import Image
    return np.array(Image.open(os.path.join(GOLDEN, name + ".png")))


def synthetic_plane(name):
    dims = {"bench_1080p": (1920, 1080), "unit_12x8": (12, 8), "unit_8x8": (8, 8)}[name]
    yy, xx = np.mgrid[0:dims[1], 0:dims[0]]
    return ((xx * yy) & 255).astype(np.uint8)   # benches/bench.rs:24-28, src/lib.rs:36-43


def get_plane(name):
    if name in ("bench_1080p", "unit_12x8", "unit_8x8"):
        return synthetic_plane(name)
    return load_plane(name)


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "golden.json")) as f:
        return json.load(f)


def sha16(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def photo_like(w, h, seed):
    """Seeded smooth-gradient + noise plane (builder-defined) so that the fix-up
    branch and the symbol mix resemble natural images."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = 128 + 90 * np.sin(xx / 37.0 + seed) * np.cos(yy / 53.0) + 30 * np.sin((xx + yy) / 11.0)
    return np.clip(base + rng.normal(0, 6, (h, w)), 0, 255).astype(np.uint8)
