"""pytest configuration: `gpu` marker + shared fixtures (golden vectors, synthetic weights)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def charset():
    return load_golden("charset")


@pytest.fixture(scope="session")
def synth_sd():
    """Synthetic reference-layout state dict (seed 0), the weights every golden logit was produced with."""
    from oracle.synth import synth_state_dict

    return synth_state_dict(0)
