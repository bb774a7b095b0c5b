"""Shared fixtures.  `-m "not gpu"` runs on the CPU-only build box, `-m gpu` on a B200."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); skipped by -m 'not gpu'")
    config.addinivalue_line("markers", "slow: minutes of host work (50 GB synthetic trajectory + CPU oracle)")


@pytest.fixture(scope="session")
def gold_si():
    return dict(np.load(GOLDEN / "si_small.npz", allow_pickle=False))


@pytest.fixture(scope="session")
def gold_gr():
    return dict(np.load(GOLDEN / "graphene_small.npz", allow_pickle=False))
