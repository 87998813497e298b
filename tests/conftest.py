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
    """Flat npz with 'case/key' names -> {case: {key: array}}."""
    z = np.load(os.path.join(GOLDEN, name))
    out = {}
    for k in z.files:
        case, _, key = k.rpartition("/")
        out.setdefault(case, {})[key] = z[k]
    return out


@pytest.fixture(scope="session")
def dcn_cases():
    return load_golden("dcn_cases.npz")


@pytest.fixture(scope="session")
def assign_cases():
    return load_golden("assign_cases.npz")


@pytest.fixture(scope="session")
def target_cases():
    """Outputs of the reference's own target-assignment / FCOSRepPoints loss code (tests/golden/gen_target_golden.py)."""
    return load_golden("target_cases.npz")


@pytest.fixture(scope="session")
def loss_cases():
    return load_golden("loss_cases.npz")


def rel_err(a, b):
    """||a-b|| / max(||b||, tiny): the relative error BASELINE.json's tolerances are stated in."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
