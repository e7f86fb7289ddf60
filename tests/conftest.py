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


def pytest_sessionfinish(session, exitstatus):
    """Largest per-step errors the GPU parity tests saw (tests/test_gpu_parity.py::within), kept next
    to the other run artefacts: the per-solver bars of DESIGN.md section 3 are read off this file."""
    try:
        import test_gpu_parity as GP
    except Exception:
        return
    if not GP.MARGINS:
        return
    out_dir = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(out_dir):
        return
    try:
        from pykrylov_b200.device import device_count
        tag = "gpu" if device_count() > 0 else "emulated"
    except Exception:
        tag = "emulated"
    with open(os.path.join(out_dir, "parity_margins_%s.json" % tag), "w") as fh:
        json.dump(GP.MARGINS, fh, indent=1, sort_keys=True)
