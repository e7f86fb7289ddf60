"""CPU tests that PIN THE ORACLE: the numpy/C restatement (oracle/) must agree with
(a) scipy's CSR kernels bit-for-bit, (b) the committed golden vectors that were
produced by the reference itself (tests/golden/make_golden.py), and (c) -- when
oracle/_ref exists (build container) -- the transliterated reference run live."""
import os
import sys
import zlib

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT, mtx
from oracle import krylov_ref as kr
from oracle.csr_ref import CsrRef, load_mtx

REF_DIR = os.path.join(ROOT, "oracle", "_ref")
HAVE_REF = os.path.isdir(os.path.join(REF_DIR, "refpykrylov"))


def ref_modules():
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import refpykrylov.linop as lo
    from refpykrylov.cg import CG
    from refpykrylov.cgs import CGS
    from refpykrylov.tfqmr import TFQMR
    from refpykrylov.bicgstab import BiCGSTAB
    from refpykrylov.minres import Minres
    return lo, dict(cg=CG, cgs=CGS, tfqmr=TFQMR, bicgstab=BiCGSTAB, minres=Minres)


# ------------------------------------------------------------ integer work / SpMV
@pytest.mark.parametrize("name", ["1138bus", "jpwh_991", "GD97_b"])
def test_fixture_csr_matches_golden_crc(name, golden):
    M = load_mtx(mtx(name))
    g = golden["csr/" + name]
    assert list(M.shape) == g["shape"] and M.nnz == g["nnz"]
    assert zlib.crc32(M.indptr.tobytes()) == g["indptr_crc"]
    assert zlib.crc32(M.indices.tobytes()) == g["indices_crc"]
    assert zlib.crc32(M.data.tobytes()) == g["data_crc"]


@pytest.mark.parametrize("name", ["1138bus", "jpwh_991", "GD97_b"])
def test_c_matvec_is_scipy_bit_for_bit(name):
    M = load_mtx(mtx(name))
    S = M.to_scipy()
    rng = np.random.default_rng(0)
    for _ in range(3):
        x = rng.standard_normal(M.shape[1])
        assert np.array_equal(M.matvec(x), S @ x)
        assert np.array_equal(M.rmatvec(x), S.T @ x)


def test_c_matvec_random_rectangular_and_empty_rows():
    R = sp.random(300, 200, density=0.02, random_state=3, format="csr")
    R.sort_indices()
    M = CsrRef.from_scipy(R)
    x = np.random.default_rng(1).standard_normal(200)
    y = np.random.default_rng(2).standard_normal(300)
    assert np.array_equal(M.matvec(x), R @ x)
    assert np.array_equal(M.rmatvec(y), R.T @ y)
    Z = CsrRef((4, 3), np.zeros(5, np.int32), np.zeros(0, np.int32), np.zeros(0))
    assert np.array_equal(Z.matvec(np.ones(3)), np.zeros(4))


def test_stencil_generators_match_scipy_kron():
    for g in (3, 10, 17):
        ip, ix, dv = kr.poisson2d_csr(g)
        T = sp.diags([-1, 2, -1], [-1, 0, 1], shape=(g, g))
        M = (sp.kron(sp.identity(g), T) + sp.kron(T, sp.identity(g))).tocsr()
        M.sort_indices()
        M.eliminate_zeros()
        assert np.array_equal(ip, M.indptr) and np.array_equal(ix, M.indices) and np.array_equal(dv, M.data)
        # row shards concatenate to the full matrix
        h = (g * g) // 2
        ip1, ix1, dv1 = kr.poisson2d_csr(g, 0, h)
        ip2, ix2, dv2 = kr.poisson2d_csr(g, h, g * g)
        assert np.array_equal(np.concatenate([ix1, ix2]), ix)
        assert np.array_equal(np.concatenate([ip1, ip2[1:] + ip1[-1]]), ip)


def test_gallery_restatement_agrees_with_csr_to_1ulp():
    g = 30
    ip, ix, dv = kr.poisson2d_csr(g)
    M = CsrRef((g * g, g * g), ip, ix, dv)
    x = np.random.default_rng(4).standard_normal(g * g)
    a, b = kr.poisson2d_matvec(x.copy()), M.matvec(x)
    assert np.allclose(a, b, rtol=0, atol=8 * np.finfo(float).eps * np.abs(x).max() * 8)


# ------------------------------------------------------------ golden vectors (from the reference)
def _check(rec, st, hist_rtol=0.0):
    assert int(st.nMatvec) == rec["nMatvec"]
    assert float(st.residNorm0) == rec["residNorm0"]
    assert float(st.residNorm) == rec["residNorm"]
    assert bool(st.converged) == rec["converged"]


def test_cg_known_answers_poisson1d(golden):
    for n in (10, 20, 100, 1000):
        e = np.ones(n)
        st = kr.cg_solve(kr.poisson1d_matvec, kr.poisson1d_matvec(e))
        rec = golden["cg/poisson1d/%d" % n]
        _check(rec, st)
        assert st.nMatvec == n // 2
        assert [float(v) for v in st.residHistory[:25]] == rec["residHistory"]


def test_cg_known_answers_poisson2d(golden):
    for g, nmv in ((10, 15), (20, 33), (100, 160)):
        e = np.ones(g * g)
        st = kr.cg_solve(kr.poisson2d_matvec, kr.poisson2d_matvec(e))
        _check(golden["cg/poisson2d_gallery/%d" % g], st)
        assert st.nMatvec == nmv
        ip, ix, dv = kr.poisson2d_csr(g)
        M = CsrRef((g * g, g * g), ip, ix, dv)
        st = kr.cg_solve(M, M.matvec(e))
        _check(golden["cg/poisson2d_csr/%d" % g], st)


def test_cg_1138bus_golden(golden):
    M = load_mtx(mtx("1138bus"))
    n = M.shape[0]
    e = np.ones(n)
    st = kr.cg_solve(M, M.matvec(e))
    _check(golden["cg/1138bus/default"], st)
    st = kr.cg_solve(M, M.matvec(e), guess=1 + np.arange(n, dtype=float), reltol=1e-8, matvec_max=2 * n)
    _check(golden["cg/1138bus/demo"], st)
    dinv = 1.0 / M.to_scipy().diagonal()
    st = kr.cg_solve(M, M.matvec(e), precon=lambda r: dinv * r)
    _check(golden["cg/1138bus/jacobi"], st)


@pytest.mark.parametrize("name,solve", [("CGS", kr.cgs_solve), ("TFQMR", kr.tfqmr_solve),
                                        ("Bi-CGSTAB", kr.bicgstab_solve)])
def test_bmark_golden(name, solve, golden, golden_vectors):
    M = load_mtx(mtx("jpwh_991"))
    n = M.shape[0]
    rhs = M.matvec(np.ones(n))
    for reltol in (1e-8, 1e-5):
        st = solve(M, rhs, guess=1 + np.arange(n, dtype=float), reltol=reltol, matvec_max=2 * n)
        key = "%s/jpwh_991/reltol%g" % (name, reltol)
        _check(golden[key], st)
        assert np.array_equal(st.x, golden_vectors[key.replace("/", "_") + "_x"])
    b = M.matvec(np.random.default_rng(5).standard_normal(n))
    _check(golden["%s/jpwh_991/zero_guess" % name], solve(M, b, reltol=1e-8, matvec_max=2 * n))


def test_bmark_doc_table_values(golden):
    """doc/source/{cgs,bicgstab}.rst:50-52 print these digits."""
    g = golden["CGS/jpwh_991/reltol1e-05"]
    assert g["nMatvec"] == 64 and "%8.2e" % g["residNorm"] == "4.72e-03" and "%8.2e" % g["err"] == "1.47e-04"
    g = golden["Bi-CGSTAB/jpwh_991/reltol1e-05"]
    assert g["nMatvec"] == 57 and "%8.2e" % g["residNorm"] == "5.18e-02" and "%8.2e" % g["err"] == "3.35e-03"
    g = golden["cg/poisson1d/100"]
    assert g["nMatvec"] == 50 and "%7.2e" % g["residNorm"] == "7.39e-14"


def test_minres_golden(golden, golden_vectors):
    M = load_mtx(mtx("jpwh_991"))
    n = M.shape[0]
    S0 = M.to_scipy()
    S = CsrRef.from_scipy((S0 + S0.T) * 0.5)
    e = np.ones(n)
    st = kr.minres_solve(S, S.matvec(e))
    g = golden["MINRES/sym_jpwh_991"]
    assert (st.istop, st.itn) == (g["istop"], g["itn"])
    for k in ("rnorm", "Anorm", "Acond", "ynorm", "Arnorm", "residNorm0"):
        assert float(st[k]) == g[k], k
    assert [float(v) for v in st.residHistory[:25]] == g["residHistory"]
    assert [float(v) for v in st.dir_errors_window[:25]] == g["dir_errors_window"]
    assert np.array_equal(st.x, golden_vectors["MINRES_sym_jpwh_991_x"])
    st = kr.minres_solve(S, S.matvec(e), shift=0.5)
    g = golden["MINRES/sym_jpwh_991_shift0.5"]
    assert (st.istop, st.itn, float(st.rnorm)) == (g["istop"], g["itn"], g["rnorm"])
    assert not kr.check_symmetric(M, n)                   # golden["MINRES/jpwh_991_nonsym"]: istop 7
    assert kr.check_symmetric(S, n)


# ------------------------------------------------------------ live reference (build container only)
@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref not generated (needs /root/reference)")
def test_port_is_bit_identical_to_live_reference():
    lo, K = ref_modules()
    M = load_mtx(mtx("jpwh_991"))
    n = M.shape[0]
    rng = np.random.default_rng(11)
    rhs = M.matvec(rng.standard_normal(n))
    guess = rng.standard_normal(n)
    op = lo.LinearOperator(n, n, lambda v: M.matvec(v))
    dvec = 1.0 / np.maximum(np.abs(M.to_scipy().diagonal()), 1.0)
    for name, solve in (("cgs", kr.cgs_solve), ("tfqmr", kr.tfqmr_solve), ("bicgstab", kr.bicgstab_solve)):
        for precon in (None, dvec):
            ks = K[name](op, reltol=1e-7, precon=None if precon is None else lo.DiagonalOperator(precon))
            ks.solve(rhs, guess=guess.copy(), matvec_max=150)
            st = solve(M, rhs, guess=guess.copy(), reltol=1e-7, matvec_max=150,
                       precon=None if precon is None else (lambda r: precon * r))
            assert (ks.nMatvec, ks.residNorm0, ks.residNorm) == (st.nMatvec, st.residNorm0, st.residNorm)
            assert np.array_equal(ks.bestSolution, st.x)
    P = load_mtx(mtx("1138bus"))
    n = P.shape[0]
    rhs = P.matvec(np.ones(n))
    op = lo.LinearOperator(n, n, lambda v: P.matvec(v), symmetric=True)
    cg = K["cg"](op)
    cg.solve(rhs, guess=guess[:1].repeat(n), matvec_max=300)
    st = kr.cg_solve(P, rhs, guess=guess[:1].repeat(n), matvec_max=300)
    assert (cg.nMatvec, cg.residNorm) == (st.nMatvec, st.residNorm)
    assert np.array_equal(cg.bestSolution, st.x) and cg.residHistory == st.residHistory
