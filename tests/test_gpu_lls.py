"""GPU tests of the solvers that run on device vectors with host-side scalars
(SURVEY.md section 8f ranks 2-3): LSQR and SYMMLQ against golden values produced by the
reference (tests/golden/golden_lls.json, generated from oracle/_ref)."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLDEN, mtx
from oracle.csr_ref import CsrRef, load_mtx

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(GOLDEN, "golden_lls.json")) as fh:
        G = json.load(fh)
    return G, np.load(os.path.join(GOLDEN, "golden_lls_vectors.npz"))


def srel(a, b):
    return abs(a - b) / abs(b) if b != 0 else abs(a)


def test_lsqr_square_and_damped(ctx, gold):
    from pykrylov_b200.linop import csr_operator
    from pykrylov_b200.lls import LSQRFramework
    G, V = gold
    M = load_mtx(mtx("jpwh_991"))
    n = M.shape[0]
    op = csr_operator(M.shape, M.indptr, M.indices, M.data, context=ctx)
    rhs = M.matvec(np.ones(n))
    for damp in (0.0, 0.1):
        ls = LSQRFramework(op, context=ctx)
        ls.solve(rhs, damp=damp, show=False)
        g = G["LSQR/jpwh_991/damp%g" % damp]
        assert ls.istop == g["istop"] and abs(ls.itn - g["itn"]) <= 3 and ls.nMatvec == 2 * ls.itn
        assert srel(ls.Anorm, g["Anorm"]) <= 2e-2 and srel(ls.xnorm, g["xnorm"]) <= 1e-6
        assert srel(ls.r2norm, g["r2norm"]) <= (0.5 if damp == 0.0 else 1e-6)    # 1e-5 residual: last-step noise
        assert np.linalg.norm(ls.x - V["LSQR_jpwh_991_damp%g_x" % damp]) <= 1e-5 * np.linalg.norm(ls.x)
        assert ls.optimal and ls.status == "direct error small"
    # self-consistency of the undamped solve: |b - A x| equals the reported r1norm
    ls = LSQRFramework(op, context=ctx)
    ls.solve(rhs, show=False, atol=1e-12, btol=1e-12, etol=0.0)
    assert abs(np.linalg.norm(rhs - M.matvec(ls.x)) - ls.r1norm) <= 1e-6 * np.linalg.norm(rhs)


def test_lsqr_rectangular_least_squares(ctx, gold, capsys):
    from pykrylov_b200.linop import linop_from_scipy
    from pykrylov_b200.lls import LSQRFramework
    G, V = gold
    R = sp.random(600, 200, density=0.03, random_state=7, format="csr")
    R.sort_indices()
    b = np.random.default_rng(7).standard_normal(600)
    op = linop_from_scipy(R, context=ctx)
    assert op.shape == (600, 200) and op.T.shape == (200, 600)
    ls = LSQRFramework(op, context=ctx)
    ls.solve(b, show=True, store_resids=True)
    g = G["LSQR/random_600x200"]
    assert (ls.istop, ls.itn) == (g["istop"], g["itn"])
    for k in ("r1norm", "r2norm", "xnorm"):
        assert srel(getattr(ls, k), g[k]) <= 1e-8, (k, getattr(ls, k), g[k])
    # Anorm / Acond accumulate every alpha_k, beta_k; after convergence the Golub-Kahan vectors
    # lose orthogonality and those late scalars are rounding-chaotic (first 8 steps agree to 1e-15)
    for k in ("Anorm", "Acond"):
        assert srel(getattr(ls, k), g[k]) <= 1e-3, (k, getattr(ls, k), g[k])
    assert np.allclose(ls.resids[:8], g["resids"][:8], rtol=1e-13)          # early steps: reduction noise only
    dmax = np.max(np.abs(np.array(ls.resids[:25]) - np.array(g["resids"])) / np.array(g["resids"]))
    assert dmax <= 1e-6, dmax
    xg = V["LSQR_random_600x200_x"]
    assert np.linalg.norm(ls.x - xg) <= 1e-7 * np.linalg.norm(xg), np.linalg.norm(ls.x - xg)
    # normal equations: A^T (b - A x) ~ 0 (the solve stops on the direct-error test, etol 1e-6)
    ne = np.linalg.norm(R.T @ (b - R @ ls.x)) / np.linalg.norm(R.T @ b)
    assert ne <= 1e-3, ne
    assert "LSQR finished" in capsys.readouterr().out


def test_lsqr_zero_rhs_and_closure_operator(ctx):
    from pykrylov_b200.linop import LinearOperator
    from pykrylov_b200.lls import LSQRFramework
    A = np.array([[2.0, 1.0], [1.0, -3.0], [0.5, 0.25]])
    op = LinearOperator(2, 3, lambda v: A @ v, matvec_transp=lambda u: A.T @ u)
    ls = LSQRFramework(op, context=ctx)
    ls.solve(np.zeros(3))
    assert ls.istop == 0 and ls.itn == 0 and np.array_equal(ls.x, np.zeros(2))
    b = np.array([1.0, 2.0, 3.0])
    ls = LSQRFramework(op, context=ctx)
    ls.solve(b, etol=0.0)
    assert np.allclose(ls.x, np.linalg.lstsq(A, b, rcond=None)[0], rtol=1e-8)


def test_symmlq_golden(ctx, gold):
    from pykrylov_b200.linop import csr_operator
    from pykrylov_b200.symmlq import Symmlq
    G, V = gold
    M = load_mtx(mtx("jpwh_991"))
    S0 = M.to_scipy()
    S = CsrRef.from_scipy((S0 + S0.T) * 0.5)
    n = S.shape[0]
    op = csr_operator(S.shape, S.indptr, S.indices, S.data, symmetric=True, context=ctx)
    rhs = S.matvec(np.ones(n))
    for shift, key in ((None, "SYMMLQ/sym_jpwh_991"), (0.5, "SYMMLQ/sym_jpwh_991_shift0.5")):
        sq = Symmlq(op, context=ctx)
        sq.solve(rhs, **({} if shift is None else {"shift": shift}))
        g = G[key]
        assert abs(sq.nMatvec - g["nMatvec"]) <= 2 and sq.converged
        assert srel(sq.anorm, g["anorm"]) <= 2e-2 and srel(sq.xNorm, g["xNorm"]) <= 1e-6
        xg = V[key.replace("/", "_") + "_x"]
        assert np.linalg.norm(sq.x - xg) <= 1e-5 * np.linalg.norm(xg)
        resid = rhs - (S.matvec(sq.x) - (shift or 0.0) * sq.x)
        assert abs(np.linalg.norm(resid) - sq.residNorm) <= 1e-8 * np.linalg.norm(rhs)
    B = load_mtx(mtx("1138bus"))
    opb = csr_operator(B.shape, B.indptr, B.indices, B.data, symmetric=True, context=ctx)
    sq = Symmlq(opb, context=ctx)
    sq.solve(B.matvec(np.ones(B.shape[0])), matvec_max=400)
    g = G["SYMMLQ/1138bus_max400"]
    # cond ~ 1e7: after 400 Lanczos steps two correct implementations differ in every digit of
    # x; only the control flow is comparable.  The reference leaves the
    # loop through `while nMatvec < matvec_max` with istop still 0 (symmlq.py:234).
    assert sq.nMatvec == g["nMatvec"] and sq.istop == 0
    assert srel(sq.anorm, g["anorm"]) <= 5e-2


def test_lsmr_rectangular_least_squares(ctx, gold):
    from pykrylov_b200.linop import linop_from_scipy
    from pykrylov_b200.lls import LSMRFramework
    G, V = gold
    R = sp.random(600, 200, density=0.03, random_state=7, format="csr")
    R.sort_indices()
    b = np.random.default_rng(7).standard_normal(600)
    ls = LSMRFramework(linop_from_scipy(R, context=ctx), context=ctx)
    x, istop, itn, normr, normar, normA, condA, normx = ls.solve(b)
    g = G["LSMR/random_600x200"]
    assert (istop, itn) == (g["istop"], g["itn"])
    assert srel(normr, g["normr"]) <= 1e-8 and srel(normx, g["normx"]) <= 1e-8
    assert srel(normA, g["normA"]) <= 1e-3 and srel(condA, g["condA"]) <= 1e-2
    xg = V["LSMR_random_600x200_x"]
    assert np.linalg.norm(x - xg) <= 1e-7 * np.linalg.norm(xg) and x is ls.x
    assert np.linalg.norm(R.T @ (b - R @ x)) <= 1e-3 * np.linalg.norm(R.T @ b)


def test_closure_operators_through_the_bridge_on_gpu(ctx):
    """Every solver accepts the reference's closure-defined operators: vectors stay in HBM,
    fused AXPY/dot kernels run on the GPU, the closure is applied on the host."""
    from oracle import krylov_ref as kr
    from pykrylov_b200.linop import LinearOperator
    from pykrylov_b200.bicgstab import BiCGSTAB
    from pykrylov_b200.cgs import CGS
    from pykrylov_b200.tfqmr import TFQMR
    from pykrylov_b200.minres import Minres
    M = load_mtx(mtx("jpwh_991"))
    n = M.shape[0]
    rng = np.random.default_rng(8)
    rhs = M.matvec(rng.standard_normal(n))
    op = LinearOperator(n, n, lambda v: M.matvec(v))
    for K, solve in ((BiCGSTAB, kr.bicgstab_solve), (CGS, kr.cgs_solve), (TFQMR, kr.tfqmr_solve)):
        ks = K(op, reltol=1e-8, context=ctx)
        ks.solve(rhs, matvec_max=2 * n)
        ref = solve(M, rhs, reltol=1e-8, matvec_max=2 * n)
        assert ks.converged and abs(ks.nMatvec - ref.nMatvec) <= 6
        assert np.linalg.norm(rhs - M.matvec(ks.bestSolution)) <= 1e-6 * np.linalg.norm(rhs)
    S0 = M.to_scipy()
    S = CsrRef.from_scipy((S0 + S0.T) * 0.5)
    d = 1.0 + rng.random(n)
    mr = Minres(LinearOperator(n, n, lambda v: S.matvec(v), symmetric=True), context=ctx)
    mr.solve(S.matvec(np.ones(n)), show=False, precon=LinearOperator(n, n, lambda v: v / d, symmetric=True))
    ref = kr.minres_solve(S, S.matvec(np.ones(n)), precon=lambda v: v / d)
    assert mr.istop == ref.istop and abs(mr.itn - ref.itn) <= 2
    assert np.linalg.norm(mr.x - ref.x) <= 1e-5 * np.linalg.norm(ref.x)


def test_craig_and_craigmr_on_gpu(ctx, capsys):
    """Least-norm solvers on a consistent under-determined system: CUDA SpMV with A and the
    device-built A^T; checked against the oracle-side NumPy bridge run of the same code
    (bit-identical to the reference on CPU, tests/test_host_solvers.py) and self-consistency."""
    import pykrylov_b200._engine as eng
    from fake_bridge import FakeBridge
    from pykrylov_b200.linop import LinearOperator, linop_from_scipy
    from pykrylov_b200.lls import CRAIGFramework, CRAIGMRFramework
    rng = np.random.default_rng(14)
    R = sp.random(90, 150, density=0.1, random_state=5, format="csr")
    R.sort_indices()
    b = R @ rng.standard_normal(150)
    cr = CRAIGFramework(linop_from_scipy(R, context=ctx), context=ctx)
    cr.solve(b)
    # (the reference stops this run on the truncated direct-error test, istop 8, with
    #  |R x - b| = 3.46 -- `optimal` says nothing about the residual; parity is what counts)
    assert cr.optimal
    cm = CRAIGMRFramework(linop_from_scipy(R, context=ctx), context=ctx)
    cm.solve(b)
    capsys.readouterr()
    real = eng.HostBridge
    eng.HostBridge = FakeBridge
    try:
        op = LinearOperator(150, 90, lambda v: R @ v, matvec_transp=lambda u: R.T @ u)
        cr_ref = CRAIGFramework(op)
        cr_ref.solve(b)
        cm_ref = CRAIGMRFramework(op)
        cm_ref.solve(b)
    finally:
        eng.HostBridge = real
    capsys.readouterr()
    assert (cr.istop, cr.itn) == (cr_ref.istop, cr_ref.itn) and abs(cr.r1norm - cr_ref.r1norm) <= 1e-9 * max(cr_ref.r1norm, 1e-30) + 1e-12
    assert np.linalg.norm(cr.x - cr_ref.x) <= 1e-8 * np.linalg.norm(cr_ref.x)
    assert (cm.istop, cm.itn) == (cm_ref.istop, cm_ref.itn)
    assert np.linalg.norm(cm.x - cm_ref.x) <= 1e-8 * np.linalg.norm(cm_ref.x)
