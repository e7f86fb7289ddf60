"""A slice of the GPU parity tests on the host emulation in its SIMT mode (tests/emu): every CUDA
thread of a block is a fiber, __syncthreads() and the warp shuffles are real barriers between
them, __shared__ variables are shared by the block, and the library's genuine
block_reduce_finalize -- warp butterflies, per-warp partials, the last-CTA ticket, the ordered
final sum and the fused scalar recurrence -- runs as written, instead of the sequential stand-in
the faster default mode uses (tests/test_emulated_device_logic.py).  ~100x slower, hence a
slice: the fused-dot / multi-AXPY entry points, determinism of the reduction, and the
single-step parity of the five loops."""
import ctypes as C

import pytest

import test_gpu_parity as GP
from test_gpu_parity import cg_form          # noqa: F401


@pytest.fixture(scope="module")
def ctx(emu_ctx):
    from pykrylov_b200 import _lib as L
    L.lib.kry_emu_set_fibers.restype = C.c_int
    L.lib.kry_emu_set_fibers.argtypes = [C.c_int]
    L.lib.kry_emu_set_fibers(1)
    try:
        yield emu_ctx
    finally:
        L.lib.kry_emu_set_fibers(0)


SLICE = ("test_spmv_fused_dots", "test_multi_axpy_dot_matches_numpy", "test_reduction_is_deterministic",
         "test_cg_single_step_from_identical_state", "test_bicgstab_single_step_from_identical_state",
         "test_cgs_single_step_from_identical_state", "test_tfqmr_single_step_from_identical_state",
         "test_minres_single_step_from_identical_state", "test_cg_edge_cases")
for _name in SLICE:
    globals()[_name + "__simt"] = getattr(GP, _name)
del _name
