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


@pytest.fixture(scope="module")
def emu_ctx():
    """A Context on the HOST EMULATION of the device logic (tests/emu, test infrastructure only):
    the library's own solvers.cu / ops.cu built for the host are swapped in behind the ctypes
    layer for the duration of the requesting module.  Skips where it cannot be built."""
    import ctypes as C
    import gc
    from emu import build_emu
    path = build_emu.build()
    if path is None:
        pytest.skip("host emulation cannot be built here (needs g++ and the CUDA headers)")
    from pykrylov_b200 import _lib as L
    from pykrylov_b200 import device
    emu = C.CDLL(path)
    for name, (restype, argtypes) in L.PROTOTYPES.items():
        fn = getattr(emu, name)
        fn.restype = restype
        fn.argtypes = argtypes
    real_lib, real_default = L.lib, device._default
    L.lib, device._default = emu, None
    c = device.Context(0)
    try:
        yield c
    finally:
        # everything created on the emulated library must be gone before the real one is back
        c.close()
        if device._default is not None:
            device._default.close()
        gc.collect()
        device.result_pool.trim()
        L.lib, device._default = real_lib, real_default
