import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def c1():
    """BASELINE.json config 1: synthetic 100K-node avg-deg-15 graph, 128-d features."""
    import legion_b200
    return legion_b200.synth.make_dataset(100_000, 15.0, 128)


@pytest.fixture(scope="session")
def small():
    import legion_b200
    return legion_b200.synth.make_dataset(5_000, 8.0, 100, n_class=7)
