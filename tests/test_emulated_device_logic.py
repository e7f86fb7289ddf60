"""The GPU parity tests, run on the CPU against a host emulation of the device logic.

tests/emu builds the library's OWN csrc/solvers.cu and csrc/ops.cu for the host (g++
-DKRY_EMULATE): the kernel loops of spmv_row_kernel / vec_pass_kernel / vec_map_kernel, every
solver functor (bodies, epilogues, scalar recurrences, stopping tests), the launch sequences of
the three CG plans, the settle logic and the whole C ABI run unchanged, with one host thread
playing every CUDA thread in turn.  That library is swapped in behind the ctypes layer for the
duration of this module, and the test functions of tests/test_gpu_parity.py and
tests/test_gpu_lls.py are collected here a second time with `ctx` bound to an emulated context.

What this does and does not show.  It checks the *logic* of the device code -- operation order
per element and per row (the file is built with -ffp-contract=off), pointer roles, buffer
rotations, flags, counters, the Python host layer on top -- against the oracle on every
machine, GPU or not.  It does not exercise the memory system, warp shuffles, atomics, PTX,
CUDA graphs, the smem / TMA kernel variants or NCCL; the `-m gpu` suite remains the parity
proof.  The emulation is test infrastructure: the product never loads it
(tests/test_host.py::test_product_never_imports_the_oracle covers tests/emu as well).
"""
import pytest

import test_gpu_lls as GL
import test_gpu_parity as GP

# fixtures the collected tests ask for
from test_gpu_parity import cg_form, minres_plan          # noqa: F401
from test_gpu_lls import gold                # noqa: F401


@pytest.fixture(scope="module")
def ctx(emu_ctx):
    """The collected tests ask for `ctx`: here it is the emulated context (conftest.emu_ctx)."""
    return emu_ctx


def _emulable(name):
    skip = ("fullsize",                   # 10^7-row operators: minutes on one host thread
            "gallery_equals_oracle",      # would test emu_context.cpp's generators, not the product
            "handles_can_be_destroyed",   # creates its own Context on cuda:0
            "one_cta_loop",               # shared memory + __syncthreads(): CUDA only
            "lsqr_rectangular")           # its 1e-3 bar on Anorm/Acond (rounding-chaotic after convergence,
                                          # see the test) is calibrated to the GPU's reduction tree
    return name.startswith("test_") and not any(s in name for s in skip)


for _mod in (GP, GL):
    for _name in dir(_mod):
        if _emulable(_name) and callable(getattr(_mod, _name)):
            globals()[_name + "__emulated"] = getattr(_mod, _name)
del _mod, _name
