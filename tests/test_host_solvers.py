"""CPU tests of the HOST-SIDE control logic of the solvers that drive device vectors
from Python (LSQR, SYMMLQ, bridged CG).  The device calls are replaced by the
test-only NumPy restatement in tests/fake_bridge.py, so what is checked here is the
scalar recurrences / stopping tests / attribute contract -- bit-for-bit against the
live reference when oracle/_ref is present, and against the golden values otherwise.
(The same solvers on real CUDA kernels are covered by tests/test_gpu_lls.py.)"""
import json
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLDEN, ROOT, mtx
from fake_bridge import FakeBridge
from oracle.csr_ref import CsrRef, load_mtx

REF_DIR = os.path.join(ROOT, "oracle", "_ref")
HAVE_REF = os.path.isdir(os.path.join(REF_DIR, "refpykrylov"))


@pytest.fixture()
def fake(monkeypatch):
    import pykrylov_b200._engine as eng
    monkeypatch.setattr(eng, "HostBridge", FakeBridge)
    return eng


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLDEN, "golden_lls.json")) as fh:
        return json.load(fh), np.load(os.path.join(GOLDEN, "golden_lls_vectors.npz"))


def sym_jpwh():
    M = load_mtx(mtx("jpwh_991"))
    S0 = M.to_scipy()
    return CsrRef.from_scipy((S0 + S0.T) * 0.5)


def test_lsqr_host_logic_matches_golden(fake, gold):
    from pykrylov_b200.linop import LinearOperator
    from pykrylov_b200.lls import LSQRFramework
    G, V = gold
    M = load_mtx(mtx("jpwh_991"))
    n = M.shape[0]
    op = LinearOperator(n, n, lambda v: M.matvec(v), matvec_transp=lambda u: M.rmatvec(u))
    for damp in (0.0, 0.1):
        ls = LSQRFramework(op)
        ls.solve(M.matvec(np.ones(n)), damp=damp)
        g = G["LSQR/jpwh_991/damp%g" % damp]
        for k in ("istop", "itn", "r1norm", "r2norm", "Anorm", "Acond", "Arnorm", "xnorm"):
            assert getattr(ls, k) == g[k], k
        assert np.array_equal(ls.x, V["LSQR_jpwh_991_damp%g_x" % damp])
        assert ls.nMatvec == 2 * ls.itn and ls.optimal
    R = sp.random(600, 200, density=0.03, random_state=7, format="csr")
    R.sort_indices()
    b = np.random.default_rng(7).standard_normal(600)
    ls = LSQRFramework(LinearOperator(200, 600, lambda v: R @ v, matvec_transp=lambda u: R.T @ u))
    ls.solve(b, store_resids=True)
    g = G["LSQR/random_600x200"]
    assert (ls.istop, ls.itn, ls.r1norm, ls.Anorm, ls.Acond) == (g["istop"], g["itn"], g["r1norm"], g["Anorm"], g["Acond"])
    assert ls.resids[:25] == g["resids"] and np.array_equal(ls.x, V["LSQR_random_600x200_x"])


def test_symmlq_host_logic_matches_golden(fake, gold):
    from pykrylov_b200.linop import LinearOperator
    from pykrylov_b200.symmlq import Symmlq
    G, V = gold
    S = sym_jpwh()
    n = S.shape[0]
    op = LinearOperator(n, n, lambda v: S.matvec(v), symmetric=True)
    for shift, key in ((None, "SYMMLQ/sym_jpwh_991"), (0.5, "SYMMLQ/sym_jpwh_991_shift0.5")):
        sq = Symmlq(op)
        sq.solve(S.matvec(np.ones(n)), **({} if shift is None else {"shift": shift}))
        g = G[key]
        assert (sq.nMatvec, sq.residNorm, sq.xNorm, sq.anorm, sq.acond) == \
            (g["nMatvec"], g["residNorm"], g["xNorm"], g["anorm"], g["acond"])
        assert np.array_equal(sq.x, V[key.replace("/", "_") + "_x"])


def test_bridged_cg_host_logic_matches_oracle(fake):
    from oracle import krylov_ref as kr
    from pykrylov_b200.linop import LinearOperator
    from pykrylov_b200.gallery import Poisson2dMatvec
    from pykrylov_b200.cg import CG
    n = 400
    A = LinearOperator(n, n, lambda x: Poisson2dMatvec(x), symmetric=True)
    rhs = A * np.ones(n)
    guess = np.random.default_rng(3).standard_normal(n)
    dinv = np.full(n, 0.25)
    from pykrylov_b200.linop import LinearOperator as LO
    P = LO(n, n, lambda r: dinv * r, symmetric=True)            # opaque preconditioner -> bridge
    for kw, okw in ((dict(), dict()), (dict(guess=guess), dict(guess=guess.copy()))):
        for precon, pfun in ((None, None), (P, lambda r: dinv * r)):
            cg = CG(A, precon=precon)
            cg.solve(rhs, store_iterates=True, **kw)
            ref = kr.cg_solve(kr.poisson2d_matvec, rhs, precon=pfun, **okw)
            assert (cg.nMatvec, cg.residNorm0, cg.residNorm, cg.converged) == \
                (ref.nMatvec, ref.residNorm0, ref.residNorm, ref.converged)
            assert cg.residHistory == ref.residHistory and np.array_equal(cg.bestSolution, ref.x)
            assert len(cg.iterates) == len(ref.residHistory)


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not generated (needs /root/reference)")
def test_lsqr_and_symmlq_bit_identical_to_live_reference(fake):
    sys.path.insert(0, REF_DIR)
    from refpykrylov.linop import LinearOperator as RLO
    from refpykrylov.lls import LSQRFramework as RLSQR
    from refpykrylov.symmlq import Symmlq as RSymmlq
    from pykrylov_b200.linop import LinearOperator
    from pykrylov_b200.lls import LSQRFramework
    from pykrylov_b200.symmlq import Symmlq
    rng = np.random.default_rng(12)
    R = sp.random(150, 90, density=0.1, random_state=3, format="csr")
    b = rng.standard_normal(150)
    mv, rmv = (lambda v: R @ v), (lambda u: R.T @ u)
    Mdiag, Ndiag = 1.0 + rng.random(150), 1.0 + rng.random(90)
    for kw in (dict(), dict(damp=0.3), dict(M=lambda u: u / Mdiag, N=lambda v: v / Ndiag), dict(atol=0, btol=0, etol=0, itnlim=60)):
        a = LSQRFramework(LinearOperator(90, 150, mv, matvec_transp=rmv))
        r = RLSQR(RLO(90, 150, mv, matvec_transp=rmv))
        a.solve(b, **kw)
        r.solve(b, **kw)
        for k in ("istop", "itn", "r1norm", "r2norm", "Anorm", "Acond", "Arnorm", "xnorm"):
            assert getattr(a, k) == getattr(r, k), (k, kw.keys())
        assert np.array_equal(a.x, r.x)
    S = sym_jpwh()
    n = S.shape[0]
    rhs = S.matvec(rng.standard_normal(n))
    for kw in (dict(), dict(shift=-0.7), dict(matvec_max=30), dict(check=True)):
        a = Symmlq(LinearOperator(n, n, lambda v: S.matvec(v), symmetric=True))
        r = RSymmlq(RLO(n, n, lambda v: S.matvec(v), symmetric=True))
        a.solve(rhs, **kw)
        r.solve(rhs, **kw)
        assert (a.nMatvec, a.residNorm, a.xNorm, a.anorm, a.acond) == (r.nMatvec, r.residNorm, r.xNorm, r.anorm, r.acond), kw
        assert np.array_equal(a.x, r.x)


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not generated (needs /root/reference)")
def test_bridged_loops_bit_identical_to_live_reference(fake, capsys):
    """Closure operators / opaque preconditioners: Bi-CGSTAB, CGS, TFQMR, MINRES host-driven
    loops (pykrylov_b200/_bridged.py) against the reference run live."""
    import io
    from contextlib import redirect_stdout
    sys.path.insert(0, REF_DIR)
    import refpykrylov.linop as rlo
    from refpykrylov.bicgstab import BiCGSTAB as RB
    from refpykrylov.cgs import CGS as RC
    from refpykrylov.tfqmr import TFQMR as RT
    from refpykrylov.minres import Minres as RM
    import pykrylov_b200.linop as lo
    from pykrylov_b200.bicgstab import BiCGSTAB
    from pykrylov_b200.cgs import CGS
    from pykrylov_b200.tfqmr import TFQMR
    from pykrylov_b200.minres import Minres
    M = load_mtx(mtx("jpwh_991"))
    n = M.shape[0]
    rng = np.random.default_rng(8)
    rhs = M.matvec(rng.standard_normal(n))
    guess = rng.standard_normal(n)
    d = 1.0 / np.maximum(np.abs(M.to_scipy().diagonal()), 1.0)
    for Ours, Ref in ((BiCGSTAB, RB), (CGS, RC), (TFQMR, RT)):
        for use_guess in (False, True):
            for use_precon in (False, True):
                a = Ours(lo.LinearOperator(n, n, lambda v: M.matvec(v)), reltol=1e-8,
                         precon=lo.LinearOperator(n, n, lambda r: d * r, symmetric=True) if use_precon else None)
                r = Ref(rlo.LinearOperator(n, n, lambda v: M.matvec(v)), reltol=1e-8,
                        precon=rlo.DiagonalOperator(d) if use_precon else None)
                kw = dict(matvec_max=2 * n)
                if use_guess:
                    a.solve(rhs, guess=guess.copy(), **kw)
                    r.solve(rhs, guess=guess.copy(), **kw)
                else:
                    a.solve(rhs, **kw)
                    r.solve(rhs, **kw)
                assert (a.nMatvec, a.residNorm0, a.residNorm, a.converged) == \
                    (r.nMatvec, r.residNorm0, r.residNorm, r.converged), (Ours.__name__, use_guess, use_precon)
                assert np.array_equal(a.bestSolution, r.bestSolution)
    S = sym_jpwh()
    b = S.matvec(np.ones(n))
    dpos = 1.0 + rng.random(n)
    for kw, pre in ((dict(), None), (dict(shift=0.3), None), (dict(), dpos), (dict(itnlim=25), None)):
        a = Minres(lo.LinearOperator(n, n, lambda v: S.matvec(v), symmetric=True))
        r = RM(rlo.LinearOperator(n, n, lambda v: S.matvec(v), symmetric=True))
        akw, rkw = dict(kw), dict(kw)
        if pre is not None:
            akw["precon"] = lo.LinearOperator(n, n, lambda v: v / pre, symmetric=True)
            rkw["precon"] = rlo.LinearOperator(n, n, lambda v: v / pre, symmetric=True)
        with redirect_stdout(io.StringIO()):
            a.solve(b, show=False, **akw)
            r.solve(b, show=False, **rkw)
        assert (a.istop, a.itn, a.rnorm, a.Anorm, a.Acond, a.ynorm, a.Arnorm) == \
            (r.istop, r.itn, r.rnorm, r.Anorm, r.Acond, r.ynorm, r.Arnorm), kw
        assert a.residHistory == r.residHistory and a.dir_errors_window == r.dir_errors_window
        assert np.array_equal(a.x, r.x)


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not generated (needs /root/reference)")
def test_lsmr_bit_identical_to_live_reference(fake):
    sys.path.insert(0, REF_DIR)
    from refpykrylov.linop import LinearOperator as RLO
    from refpykrylov.lls import LSMRFramework as RLSMR
    from pykrylov_b200.linop import LinearOperator
    from pykrylov_b200.lls import LSMRFramework
    rng = np.random.default_rng(13)
    R = sp.random(150, 90, density=0.1, random_state=3, format="csr")
    b = rng.standard_normal(150)
    mv, rmv = (lambda v: R @ v), (lambda u: R.T @ u)
    Mdiag, Ndiag = 1.0 + rng.random(150), 1.0 + rng.random(90)
    for kw in (dict(), dict(damp=0.3), dict(M=lambda u: u / Mdiag, N=lambda v: v / Ndiag),
               dict(atol=0, btol=0, etol=0, itnlim=40), dict(store_resids=True)):
        a = LSMRFramework(LinearOperator(90, 150, mv, matvec_transp=rmv))
        r = RLSMR(RLO(90, 150, mv, matvec_transp=rmv))
        ra = a.solve(b, **kw)
        rr = r.solve(b, **kw)
        assert tuple(ra[1:]) == tuple(rr[1:]), kw.keys()
        assert np.array_equal(ra[0], rr[0]) and np.array_equal(a.x, r.x)
        assert a.resids == r.resids and a.dir_errors_window == r.dir_errors_window
    M = load_mtx(mtx("jpwh_991"))
    n = M.shape[0]
    a = LSMRFramework(LinearOperator(n, n, lambda v: M.matvec(v), matvec_transp=lambda u: M.rmatvec(u)))
    r = RLSMR(RLO(n, n, lambda v: M.matvec(v), matvec_transp=lambda u: M.rmatvec(u)))
    ra, rr = a.solve(M.matvec(np.ones(n))), r.solve(M.matvec(np.ones(n)))
    assert tuple(ra[1:]) == tuple(rr[1:]) and np.array_equal(ra[0], rr[0])
    z = LSMRFramework(LinearOperator(90, 150, mv, matvec_transp=rmv)).solve(np.zeros(150))
    assert z[1] == 0 and z[2] == 0 and np.array_equal(z[0], np.zeros(90))


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not generated (needs /root/reference)")
def test_craig_and_craigmr_bit_identical_to_live_reference(fake, capsys):
    sys.path.insert(0, REF_DIR)
    from refpykrylov.linop import LinearOperator as RLO
    from refpykrylov.lls import CRAIGFramework as RC, CRAIGMRFramework as RCM
    from pykrylov_b200.linop import LinearOperator
    from pykrylov_b200.lls import CRAIGFramework, CRAIGMRFramework
    rng = np.random.default_rng(14)
    R = sp.random(90, 150, density=0.1, random_state=5, format="csr")     # m < n: consistent system
    b = R @ rng.standard_normal(150)
    mv, rmv = (lambda v: R @ v), (lambda u: R.T @ u)
    Md, Nd = 1.0 + rng.random(90), 1.0 + rng.random(150)
    for kw in (dict(), dict(M=lambda u: u / Md, N=lambda v: v / Nd), dict(itnlim=12, store_resids=True),
               dict(btol=0, etol=0, itnlim=40)):
        a = CRAIGFramework(LinearOperator(150, 90, mv, matvec_transp=rmv))
        r = RC(RLO(150, 90, mv, matvec_transp=rmv))
        a.solve(b, **kw)
        r.solve(b, **kw)
        for k in ("istop", "itn", "r1norm", "r2norm", "Arnorm", "xnorm", "nMatvec", "optimal"):
            assert getattr(a, k) == getattr(r, k), (k, list(kw))
        assert np.array_equal(a.x, r.x) and np.array_equal(a.r, r.r)
        assert a.resids == r.resids and a.dir_errors_d_window == r.dir_errors_d_window
    for kw in (dict(), dict(M=lambda u: u / Md, N=lambda v: v / Nd), dict(itnlim=12, store_resids=True)):
        a = CRAIGMRFramework(LinearOperator(150, 90, mv, matvec_transp=rmv))
        r = RCM(RLO(150, 90, mv, matvec_transp=rmv))
        a.solve(b, **kw)
        out_a = capsys.readouterr().out
        r.solve(b, **kw)
        out_r = capsys.readouterr().out
        assert (a.istop, a.itn, a.nMatvec, a.optimal) == (r.istop, r.itn, r.nMatvec, r.optimal)
        assert np.array_equal(a.x, r.x) and a.norms == r.norms and a.dir_errors_window == r.dir_errors_window
        assert out_a == out_r and out_a.startswith("1 ")            # the unconditional per-iteration print
