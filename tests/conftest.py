import json
import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"
REFERENCE = Path("/root/reference")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: reads the read-only reference mount (skipped where absent)")


@pytest.fixture(scope="session")
def golden_images():
    return dict(np.load(GOLDEN / "images.npz"))


@pytest.fixture(scope="session")
def corpus_manifest():
    return json.loads((GOLDEN / "corpus_manifest.json").read_text())


@pytest.fixture(scope="session")
def published_sizes():
    return json.loads((GOLDEN / "published_sizes.json").read_text())


def gnat_image(width, height, sigma=3.0, seed=2, phase=0):
    """SURVEY.md 8(d) config 2 generator 'G-nat' (host-generated, stored as u8)."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:height, 0:width].astype(np.float64)
    x = x + phase
    v = 128 + 60 * np.sin(x / 97) * np.cos(y / 131) + 40 * np.sin((x + y) / 37) + rng.normal(0, sigma, (height, width))
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def gnat_rgb(width, height):
    """three G-nat planes, phases (0, 11, 23), seeds 3, 4, 5."""
    return np.stack([gnat_image(width, height, seed=s, phase=p) for s, p in ((3, 0), (4, 11), (5, 23))], axis=-1)
