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
from test_gpu_parity import cg_form, minres_plan          # noqa: F401


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


def test_cg_one_cta_kernel_to_convergence__simt(ctx):
    """KRY_OPT_CG_ONE_CTA (candidate): the whole CG loop inside one CTA -- shared-memory CSR and
    vectors, __syncthreads() between the phases, block-level reductions -- played by 1024 fibers,
    to convergence on a small Laplacian, with and without a preconditioner, against the oracle."""
    import numpy as np
    from oracle import krylov_ref as kr
    from oracle.csr_ref import CsrRef
    from pykrylov_b200 import _lib as L
    from pykrylov_b200 import device as dev
    g = 9
    n = g * g
    ip, ix, dv = kr.poisson2d_csr(g)
    M = CsrRef((n, n), ip, ix, dv)
    rhs = M.matvec(np.arange(1.0, n + 1.0))
    saved = ctx.get_option(L.KRY_OPT_CG_ONE_CTA)
    ctx.set_option(L.KRY_OPT_CG_ONE_CTA, 1)
    try:
        for pmode in (0, 2):
            d = np.full(n, 4.0)
            A = dev.DeviceCsr.from_arrays(ctx, M.shape, M.indptr, M.indices, M.data, symmetric=True)
            S = dev.DeviceSolver(ctx, "cg", A)
            S.set_precon_diag(d if pmode else None, pmode)
            S.setup(rhs, matvec_max=2 * n)
            l0 = ctx.launch_count()
            S.iterate(5)
            assert ctx.launch_count() - l0 == 1                     # one launch for the whole chunk
            st = S.run(7)
            ref = kr.cg_solve(M, rhs, precon=(lambda r: r / d) if pmode else None)
            hist = S.drain_history(st)[:, 0]
            assert st.n_matvec == ref.nMatvec and bool(st.converged) == bool(ref.converged)
            rh = np.array(ref.residHistory)
            assert len(hist) == len(rh) and np.max(np.abs(hist - rh) / rh) <= 1e-9
            assert np.max(np.abs(S.solution() - ref.x)) <= 1e-10 * np.max(np.abs(ref.x))
            S._release()
            A._release()
    finally:
        ctx.set_option(L.KRY_OPT_CG_ONE_CTA, saved)
