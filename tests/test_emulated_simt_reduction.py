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


def test_minres_persistent_kernel_follows_the_oracle_and_the_3_launch_plan__simt(ctx):
    """KRY_OPT_MINRES_PERSISTENT (candidate): one cooperative kernel per iterate call -- a resident
    CTA wave, the reductions carried by grid-wide barriers -- played by the SIMT emulation's
    cooperative launch (all blocks alive at once).  To convergence and with a tiny itnlim,
    against the oracle; against the 3-launch plan to rounding; one launch per call."""
    import numpy as np
    import scipy.sparse as sp
    from oracle import krylov_ref as kr
    from oracle.csr_ref import CsrRef
    from pykrylov_b200 import _lib as L
    from pykrylov_b200 import device as dev
    rng = np.random.default_rng(8)
    saved = ctx.get_option(L.KRY_OPT_MINRES_PERSISTENT)
    try:
        for n, itnlim, shift in ((700, None, 0.0), (333, 7, 0.25), (64, None, 0.0), (257, 1, 0.0)):
            B = sp.random(n, n, density=min(1.0, 5.0 / n), random_state=int(rng.integers(1 << 30)), format="csr")
            A0 = ((B + B.T) * 0.5 + sp.diags(rng.choice([-1.0, 1.0], size=n) * (2.0 + abs(B).sum(axis=1).max()))).tocsr()
            A0.sort_indices()
            M = CsrRef.from_scipy(A0)
            rhs = M.matvec(rng.standard_normal(n))
            lim = 5 * n if itnlim is None else itnlim
            ref = kr.minres_solve(M, rhs, shift=shift, itnlim=lim)
            res = {}
            for persistent in (1, 0):
                ctx.set_option(L.KRY_OPT_MINRES_PERSISTENT, persistent)
                ctx.set_option(L.KRY_OPT_MINRES_FUSE, 0)
                A = dev.DeviceCsr.from_arrays(ctx, M.shape, M.indptr, M.indices, M.data, symmetric=True)
                S = dev.DeviceSolver(ctx, "minres", A)
                S.setup(rhs, abstol=0.0, reltol=0.0, matvec_max=lim, shift=shift, rtol=1e-12, etol=1e-6, window=5)
                l0 = ctx.launch_count()
                S.iterate(3)
                if persistent:
                    assert ctx.launch_count() - l0 == 1
                w3 = S.get_vector("w")
                st = S.run(4)
                hist = S.drain_history(st)[:, 0]
                assert (int(st.istop), int(st.n_iter)) == (ref.istop, ref.itn), (n, persistent)
                rh = np.array(ref.residHistory, dtype=float)
                k = min(len(rh), 8)
                assert len(hist) == len(rh) and np.max(np.abs(hist[:k] - rh[:k]) / rh[:k]) <= 1e-9
                xs = S.solution()
                assert np.max(np.abs(xs - ref.x)) <= 1e-7 * max(np.max(np.abs(ref.x)), 1e-300)
                res[persistent] = (hist, xs, w3, st.resid_norm)
                S._release()
                A._release()
            # the two plans sum their inner products over different grids: equal to rounding early
            # on (w after 3 trips, the first history entries), to the solver's accuracy at the end
            (h1, x1, w1, _), (h0, x0, w0, _) = res[1], res[0]
            assert np.max(np.abs(w1 - w0)) <= 1e-11 * max(np.max(np.abs(w0)), 1e-300)
            k = min(len(h0), 8)
            assert np.max(np.abs(h1[:k] - h0[:k])) <= 1e-11 * np.max(np.abs(h0[:k]))
            assert np.max(np.abs(x1 - x0)) <= 1e-7 * max(np.max(np.abs(x0)), 1e-300)
    finally:
        ctx.set_option(L.KRY_OPT_MINRES_PERSISTENT, saved)
