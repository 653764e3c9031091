import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def rel_fro(a, b):
    """||a-b||_F / ||b||_F, the parity metric of BASELINE.json (tolerance 1e-10)."""
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


@pytest.fixture(scope="session")
def golden_two_qubit():
    return dict(np.load(os.path.join(GOLDEN, "two_qubit.npz")))


@pytest.fixture(scope="session")
def golden_transmon():
    return dict(np.load(os.path.join(GOLDEN, "transmon_expanded.npz")))


@pytest.fixture(scope="session")
def golden_tf_utils():
    return dict(np.load(os.path.join(GOLDEN, "tf_utils.npz")))


@pytest.fixture(scope="session")
def golden_tunable_coupler():
    return dict(np.load(os.path.join(GOLDEN, "tunable_coupler.npz")))


@pytest.fixture(scope="session")
def golden_generator():
    return dict(np.load(os.path.join(GOLDEN, "generator.npz")))


@pytest.fixture(scope="session")
def golden_tc_levels():
    return dict(np.load(os.path.join(GOLDEN, "tunable_coupler_levels.npz")))
