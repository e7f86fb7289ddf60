import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with `-m gpu` on a B200)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "golden.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def golden_vectors():
    return np.load(os.path.join(GOLDEN, "golden_vectors.npz"))


@pytest.fixture(scope="session")
def ctx():
    """One device context for the whole GPU session (fails loudly without a GPU)."""
    from pykrylov_b200.device import Context
    c = Context(0)
    yield c
    c.close()


def mtx(name):
    return os.path.join(GOLDEN, name + ".mtx")
