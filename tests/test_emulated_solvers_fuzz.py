"""Random small systems through the five device-resident loops (on the host emulation of the
device logic, tests/emu) against the oracle's restatement of the reference loops: same
iteration counts, same convergence flags, histories and solutions to rounding.  The fixtures of
tests/test_gpu_parity.py pin a handful of matrices; this sweeps shapes the fixtures do not have
(1 x 1, empty rows, zero right-hand sides, tiny matvec_max, guesses, both diagonal
preconditioner forms) through the same C ABI."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import krylov_ref as kr
from oracle.csr_ref import CsrRef


def _system(rng, n, symmetric):
    B = sp.random(n, n, density=min(1.0, 5.0 / max(n, 1)), random_state=int(rng.integers(1 << 30)), format="csr")
    if symmetric:
        B = (B + B.T) * 0.5
    A = (B + sp.identity(n) * (2.0 + abs(B).sum(axis=1).max())).tocsr()      # diagonally dominant
    A.sort_indices()
    return CsrRef.from_scipy(A)


def _close(a, b, rtol):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    scale = max(np.max(np.abs(b)) if b.size else 0.0, 1e-300)
    return a.shape == b.shape and (a.size == 0 or np.max(np.abs(a - b)) <= rtol * scale)


CASES = [("cg", kr.cg_solve, True), ("bicgstab", kr.bicgstab_solve, False), ("cgs", kr.cgs_solve, False),
         ("tfqmr", kr.tfqmr_solve, False)]


@pytest.mark.parametrize("method,oracle,symmetric", CASES)
@pytest.mark.parametrize("seed", range(16))
def test_device_loops_follow_the_oracle_on_random_systems(emu_ctx, method, oracle, symmetric, seed):
    from pykrylov_b200 import device as dev
    rng = np.random.default_rng(77 * seed + len(method))
    n = int(rng.integers(1, 70))
    M = _system(rng, n, symmetric)
    rhs = M.matvec(rng.standard_normal(n)) if seed % 5 else np.zeros(n)
    guess = rng.standard_normal(n) if seed % 3 == 0 else None
    pmode = int(rng.integers(0, 3))
    d = np.abs(M.to_scipy().diagonal())
    pvec = None if pmode == 0 else (1.0 / d if pmode == 1 else d)
    pfun = None if pmode == 0 else ((lambda r: pvec * r) if pmode == 1 else (lambda r: r / pvec))
    matvec_max = int(rng.integers(1, 3 * n + 3)) if seed % 4 == 1 else 2 * n + 10
    if method == "tfqmr" and seed % 5 == 0:
        rhs = M.matvec(rng.standard_normal(n))      # converged-at-entry TFQMR is a NameError in the reference
    ref = oracle(M, rhs, guess=None if guess is None else guess.copy(), precon=pfun, matvec_max=matvec_max)
    A = dev.DeviceCsr.from_arrays(emu_ctx, M.shape, M.indptr, M.indices, M.data, symmetric=symmetric)
    S = dev.DeviceSolver(emu_ctx, method, A)
    S.set_precon_diag(pvec, pmode)
    S.setup(rhs, guess=guess, matvec_max=matvec_max)
    st = S.run(int(rng.integers(1, 12)))
    hist = S.drain_history(st)[:, 0]
    x = S.solution()
    S._release()
    A._release()
    assert st.n_matvec == ref.nMatvec, (st.n_matvec, ref.nMatvec)
    assert bool(st.converged) == bool(ref.converged)
    rh = np.array(ref.residHistory, dtype=float)
    assert len(hist) == len(rh)
    # leading part of the history: before rounding differences of the inner products are amplified
    k = min(len(rh), 8)
    assert _close(hist[:k], rh[:k], 1e-9), (hist[:k], rh[:k])
    assert _close(x, ref.x, 1e-7)


@pytest.mark.parametrize("seed", range(40))
def test_device_minres_follows_the_oracle_on_random_systems(emu_ctx, seed):
    from pykrylov_b200 import device as dev
    rng = np.random.default_rng(900 + seed)
    n = int(rng.integers(2, 70))
    B = sp.random(n, n, density=min(1.0, 5.0 / n), random_state=int(rng.integers(1 << 30)), format="csr")
    A0 = ((B + B.T) * 0.5 + sp.diags(rng.choice([-1.0, 1.0], size=n) * (2.0 + abs(B).sum(axis=1).max()))).tocsr()
    A0.sort_indices()
    M = CsrRef.from_scipy(A0)
    rhs = M.matvec(rng.standard_normal(n)) if seed % 5 else np.zeros(n)
    shift = float(rng.choice([0.0, 0.25]))
    itnlim = int(rng.integers(1, 2 * n)) if seed % 4 == 1 else 5 * n
    pmode = int(rng.integers(0, 3))                      # none / y = d .* r / y = r ./ d
    d = 0.5 + rng.random(n)
    pfun = None if pmode == 0 else ((lambda r: d * r) if pmode == 1 else (lambda r: r / d))
    ref = kr.minres_solve(M, rhs, precon=pfun, shift=shift, itnlim=itnlim)
    A = dev.DeviceCsr.from_arrays(emu_ctx, M.shape, M.indptr, M.indices, M.data, symmetric=True)
    S = dev.DeviceSolver(emu_ctx, "minres", A)
    S.set_precon_diag(d if pmode else None, pmode)
    S.setup(rhs, abstol=0.0, reltol=0.0, matvec_max=itnlim, shift=shift, rtol=1e-12, etol=1e-6, window=5)
    st = S.run(int(rng.integers(1, 12)))
    hist = S.drain_history(st)[:, 0]
    x = S.solution()
    S._release()
    A._release()
    assert (int(st.istop), int(st.n_iter)) == (ref.istop, ref.itn), (st.istop, st.n_iter, ref.istop, ref.itn)
    rh = np.array(ref.residHistory, dtype=float)
    assert len(hist) == len(rh) and _close(hist[:8], rh[:8], 1e-9)
    assert _close(x, ref.x, 1e-7)


def test_minres_public_api_routes_diagonal_preconditioners_to_the_device(emu_ctx, capsys):
    """Minres.solve(precon=DiagonalOperator(d)) and the bmark-style `r / diag` object both iterate
    on the device (no bridge), and agree with the reference loop; an opaque preconditioner still
    goes through the host-callback bridge and gives the same answer."""
    import pykrylov_b200._engine as eng
    from pykrylov_b200.linop import DiagonalOperator, LinearOperator, csr_operator
    from pykrylov_b200.minres import Minres
    rng = np.random.default_rng(11)
    n = 60
    B = sp.random(n, n, density=0.1, random_state=3, format="csr")
    A0 = ((B + B.T) * 0.5 + sp.diags(rng.choice([-1.0, 1.0], size=n) * 3.0)).tocsr()
    A0.sort_indices()
    M = CsrRef.from_scipy(A0)
    op = csr_operator(M.shape, M.indptr, M.indices, M.data, symmetric=True, context=emu_ctx)
    rhs = M.matvec(rng.standard_normal(n))
    d = 0.5 + rng.random(n)
    ref = kr.minres_solve(M, rhs, precon=lambda r: d * r)

    class DiagonalPrec(object):                     # examples/bmark.py:14-22
        def __init__(self, diag):
            self.diag = diag

        def __call__(self, y):
            return y / self.diag

        __mul__ = __call__

    bridged = []
    real = eng.HostBridge

    class Spy(real):
        def __init__(self, *a, **k):
            bridged.append(1)
            real.__init__(self, *a, **k)

    eng.HostBridge = Spy
    try:
        for precon in (DiagonalOperator(d), DiagonalPrec(1.0 / d)):
            mr = Minres(op, context=emu_ctx)
            mr.solve(rhs, precon=precon, show=False, check=False)
            assert not bridged
            assert (mr.istop, mr.itn) == (ref.istop, ref.itn)
            assert np.max(np.abs(mr.x - ref.x)) <= 1e-9 * np.max(np.abs(ref.x))
        mr = Minres(op, context=emu_ctx)
        mr.solve(rhs, precon=LinearOperator(n, n, lambda v: d * v, symmetric=True), show=False, check=False)
        assert bridged and (mr.istop, mr.itn) == (ref.istop, ref.itn)
        assert np.max(np.abs(mr.x - ref.x)) <= 1e-9 * np.max(np.abs(ref.x))
    finally:
        eng.HostBridge = real
    capsys.readouterr()
