"""Device-resident scalar planes of LSQR / LSMR / CRAIG / CRAIG-MR / SYMMLQ (csrc/lls.cu) on the
host emulation of the device logic: the step kernels -- the library's own lls.cu built for the
host -- must reproduce the host-driven scalar path of the same solver BIT FOR BIT (that path is
pinned bit-for-bit to the live reference in tests/test_host_solvers.py), for any check interval,
while the host synchronises only once per check interval instead of three times per trip."""
import io
from contextlib import redirect_stdout

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import mtx
from oracle.csr_ref import CsrRef, load_mtx


def _solve(K, op, b, ctx, eng, device_plane, interval, **kw):
    saved = eng.plane_csr
    if not device_plane:
        eng.plane_csr = lambda op_: None
    try:
        k = K(op, context=ctx, check_interval=interval)
        with redirect_stdout(io.StringIO()):
            ret = k.solve(b, **kw)
    finally:
        eng.plane_csr = saved
    return k, ret


CASES = [("lsqr", dict()), ("lsqr", dict(damp=0.3)), ("lsqr", dict(atol=0.0, btol=0.0, etol=0.0, itnlim=23)),
         ("lsmr", dict()), ("lsmr", dict(damp=0.2)), ("craig", dict()), ("craigmr", dict())]


@pytest.mark.parametrize("name,kw", CASES)
def test_scalar_plane_on_device_equals_host_path(emu_ctx, name, kw, monkeypatch):
    # Bit-for-bit equality needs the same launches on both sides: the host-scalar path enqueues the
    # product, the vector pass and reads the inner product; KRY_LLS_FUSE=0 makes the device plane enqueue
    # exactly those (the default fuses them into the SpMV launch, whose inner product is summed over a
    # different grid: equal to rounding, tests/test_gpu_lls.py::test_fused_trip_launches_equal_...).
    monkeypatch.setenv("KRY_LLS_FUSE", "0")
    import pykrylov_b200._engine as eng
    from pykrylov_b200 import _lib as L
    from pykrylov_b200.linop import linop_from_scipy
    from pykrylov_b200.lls import LSQRFramework, LSMRFramework, CRAIGFramework, CRAIGMRFramework
    K = dict(lsqr=LSQRFramework, lsmr=LSMRFramework, craig=CRAIGFramework, craigmr=CRAIGMRFramework)[name]
    rng = np.random.default_rng(14)
    if name in ("craig", "craigmr"):
        R = sp.random(90, 150, density=0.1, random_state=5, format="csr")
        b = R @ rng.standard_normal(150)
    else:
        R = sp.random(600, 200, density=0.03, random_state=7, format="csr")
        b = np.random.default_rng(7).standard_normal(600)
    R.sort_indices()
    op = linop_from_scipy(R, context=emu_ctx)
    host, rh = _solve(K, op, b, emu_ctx, eng, False, 32, store_resids=True, **kw)
    syncs = []
    real = L.lib.kry_lls_status

    def counting(*a):
        syncs.append(1)
        return real(*a)
    for interval in (1, 7, 32):
        del syncs[:]
        L.lib.kry_lls_status = counting
        try:
            devp, rd = _solve(K, op, b, emu_ctx, eng, True, interval, store_resids=True, **kw)
        finally:
            L.lib.kry_lls_status = real
        assert np.array_equal(devp.x, host.x), interval
        for attr in ("resids", "normal_eqns_resids", "norms"):
            if hasattr(host, attr):
                assert getattr(devp, attr) == getattr(host, attr), (attr, interval)
        for attr in ("istop", "itn", "nMatvec", "r1norm", "r2norm", "Anorm", "Acond", "Arnorm", "xnorm"):
            if hasattr(host, attr) and name != "lsmr":
                assert getattr(devp, attr) == getattr(host, attr), (attr, interval)
        if rh is not None:                                    # LSMR returns its results as a tuple
            assert all(np.array_equal(a, c) for a, c in zip(rd, rh))
        itn = devp.itn if name != "lsmr" else rd[2]
        assert 1 <= len(syncs) <= itn // interval + 3, (len(syncs), itn, interval)
    # truncated direct-error estimates: np.linalg.norm of the window on the host, an ordered sum of
    # squares on the device -- equal to rounding
    for attr in ("dir_errors_window", "dir_errors_d_window"):
        a, h = np.array(getattr(devp, attr, [])), np.array(getattr(host, attr, []))
        assert len(a) == len(h) and (len(a) == 0 or np.max(np.abs(a - h) / np.abs(h)) <= 1e-15)


@pytest.mark.parametrize("kw", [dict(), dict(shift=0.5), dict(matvec_max=40), dict(rtol=1e-14)])
def test_symmlq_scalar_plane_on_device_equals_host_path(emu_ctx, kw, monkeypatch):
    monkeypatch.setenv("KRY_LLS_FUSE", "0")               # same launches on both sides: see above
    import pykrylov_b200._engine as eng
    from pykrylov_b200.linop import csr_operator
    from pykrylov_b200.symmlq import Symmlq
    M = load_mtx(mtx("jpwh_991"))
    S0 = M.to_scipy()
    S = CsrRef.from_scipy((S0 + S0.T) * 0.5)
    n = S.shape[0]
    op = csr_operator(S.shape, S.indptr, S.indices, S.data, symmetric=True, context=emu_ctx)
    rhs = S.matvec(np.ones(n))
    host, _ = _solve(Symmlq, op, rhs, emu_ctx, eng, False, 32, **kw)
    for interval in (1, 5, 64):
        devp, _ = _solve(Symmlq, op, rhs, emu_ctx, eng, True, interval, **kw)
        assert np.array_equal(devp.x, host.x)
        for attr in ("nMatvec", "istop", "itn", "anorm", "acond", "xNorm", "residNorm", "converged"):
            assert getattr(devp, attr) == getattr(host, attr), (attr, interval)


def test_stand_alone_launches_run_again_after_a_plane_solve(emu_ctx):
    """The plane's `done` flag gates the context's stand-alone launches only while a plane loop is
    running: afterwards SpMV and fused vector ops execute unconditionally again."""
    from pykrylov_b200.linop import linop_from_scipy
    from pykrylov_b200.lls import LSQRFramework
    R = sp.random(60, 20, density=0.2, random_state=1, format="csr")
    R.sort_indices()
    op = linop_from_scipy(R, context=emu_ctx)
    b = np.random.default_rng(2).standard_normal(60)
    LSQRFramework(op, context=emu_ctx).solve(b)
    x = np.random.default_rng(3).standard_normal(20)
    assert np.array_equal(op * x, CsrRef.from_scipy(R).matvec(x))


def test_failed_trip_capture_falls_back_to_enqueued_launches(emu_ctx, monkeypatch, caplog):
    """If the one-off capture of a trip fails, nothing of it has executed: the loop keeps enqueueing the
    same launches (no other code path) and says so once."""
    from pykrylov_b200 import _lib as L
    from pykrylov_b200 import device
    from pykrylov_b200.linop import linop_from_scipy
    from pykrylov_b200.lls import LSQRFramework
    R = sp.random(300, 120, density=0.05, random_state=3, format="csr")
    R.sort_indices()
    b = np.random.default_rng(3).standard_normal(300)
    op = linop_from_scipy(R, context=emu_ctx)
    plain = LSQRFramework(op, context=emu_ctx)
    plain.solve(b, show=False, store_resids=True, check_interval=4)
    calls = []

    def failing(ctx, fn):
        calls.append(1)
        raise L.KrylovDeviceError(L.KRY_ERR_CUDA, "capture failed (injected)")
    monkeypatch.setattr(device.LaunchGraph, "capture", classmethod(lambda cls, ctx, fn: failing(ctx, fn)))
    real_get = emu_ctx.get_option                         # the emulation always answers 0 for KRY_OPT_GRAPHS
    monkeypatch.setattr(emu_ctx, "get_option", lambda o: 1 if o == L.KRY_OPT_GRAPHS else real_get(o), raising=False)
    k = LSQRFramework(op, context=emu_ctx, check_interval=4)
    with caplog.at_level("WARNING", logger="pykrylov_b200"):
        k.solve(b, show=False, store_resids=True)
    assert calls == [1] and any("capture failed" in r.getMessage() for r in caplog.records)
    assert k.itn == plain.itn and k.istop == plain.istop and k.resids == plain.resids
    assert np.array_equal(k.x, plain.x)


@pytest.mark.parametrize("seed", range(12))
def test_fused_trip_forms_agree_on_random_rectangular_systems(emu_ctx, seed, monkeypatch):
    """The three ways a trip can be enqueued -- phase of the recurrence inside the fused SpMV launch
    (KRY_LLS_FUSE=2), as its own launch behind it (3), and the three-launch form (0) -- on random sparse
    m x n systems (m < n and m > n, damped and not, random check intervals).  Forms 2 and 3 run the same
    kernels on the same data and must agree bit for bit; against form 0 only the summation order of the
    inner products differs.  Four trips: these random operators amplify a rounding-level difference by
    ~1e3 every three trips (the same growth separates either form from a long-double LSQR)."""
    import contextlib
    import io
    from pykrylov_b200.linop import linop_from_scipy
    from pykrylov_b200.lls import LSQRFramework, LSMRFramework, CRAIGFramework, CRAIGMRFramework
    rng = np.random.default_rng(seed)
    m, n = int(rng.integers(20, 400)), int(rng.integers(20, 400))
    R = sp.random(m, n, density=float(rng.uniform(0.02, 0.2)), random_state=seed, format="csr")
    R.sort_indices()
    for cls in (LSQRFramework, LSMRFramework, CRAIGFramework, CRAIGMRFramework):
        consistent = cls in (CRAIGFramework, CRAIGMRFramework)
        b = R @ rng.standard_normal(n) if consistent else rng.standard_normal(m)
        damp = 0.0 if consistent else float(rng.choice([0.0, 0.0, 0.1, 1.0]))
        interval = int(rng.integers(1, 9))
        res = {}
        for fuse in ("2", "3", "0"):
            monkeypatch.setenv("KRY_LLS_FUSE", fuse)
            k = cls(linop_from_scipy(R, context=emu_ctx), context=emu_ctx, check_interval=interval)
            with contextlib.redirect_stdout(io.StringIO()):
                k.solve(b, damp=damp, show=False, store_resids=True, itnlim=4, atol=0.0, btol=0.0, etol=0.0)
            res[fuse] = (np.array(k.resids if len(k.resids) else k.normal_eqns_resids), np.array(k.x))
        assert np.array_equal(res["2"][0], res["3"][0]) and np.array_equal(res["2"][1], res["3"][1]), cls.__name__
        h, x = res["2"]
        h0, x0 = res["0"]
        assert len(h) == len(h0) >= 4, (cls.__name__, len(h), len(h0))     # some classes record the initial residual too
        assert np.allclose(h, h0, rtol=1e-9, atol=0.0), cls.__name__
        assert np.linalg.norm(x - x0) <= 1e-9 * max(np.linalg.norm(x0), 1e-300), cls.__name__
