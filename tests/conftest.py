"""Shared test plumbing.  ``-m "not gpu"`` runs here on CPU; ``-m gpu`` runs on a B200."""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def load_golden(name: str):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def golden():
    return load_golden


@pytest.fixture()
def cpu_engine():
    """Install the oracle-backed stand-in engine (host-logic tests only)."""
    from pyparrm_b200 import _engine
    from tests.oracle_engine import OracleEngine

    engine = OracleEngine()
    _engine.set_engine(engine)
    yield engine
    _engine.set_engine(None)


@pytest.fixture(scope="session")
def gpu_engine():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("gpu-marked test started without a CUDA device")
    from pyparrm_b200 import _engine

    _engine.set_engine(None)
    return _engine.get_engine()
