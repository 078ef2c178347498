"""pytest configuration: `gpu` marker and import paths.

`-m "not gpu"` runs on the CPU-only build container; `-m gpu` needs a B200 and calls the CUDA
path through the C ABI (libauvrrt.so).
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "auv-sim_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 GPU (run with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def catalina_map():
    with open(os.path.join(GOLDEN, "catalina_map.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def shark_grid():
    z = np.load(os.path.join(GOLDEN, "shark_grid.npz"))
    return z["bins"], z["probs"]


@pytest.fixture(scope="session")
def exploring_golden():
    z = np.load(os.path.join(GOLDEN, "exploring.npz"))
    meta = json.loads(str(z["meta"]))
    return z, meta
