#!/usr/bin/env python3
"""Generate tests/golden/golden.json + golden_vectors.npz FROM THE REFERENCE ITSELF.

Runs the transliterated reference (oracle/_ref/refpykrylov, produced by
oracle/make_ref.py from /root/reference) through its own public API with the
stand-in operator of SURVEY.md section 8c (scipy CSR, sorted int32 indices) on the
shipped fixtures and the gallery problems, and records

  * the scalar outcomes the reference's docs/tests pin (nMatvec, residNorm0,
    residNorm, error) -- BASELINE.md section 1 / SURVEY.md section 4;
  * the first residHistory entries and the final iterate of each run;
  * integer work: scipy's indptr/indices for every fixture (CRC32).

This script only runs in the build container (it needs oracle/_ref, which is
generated from /root/reference); the files it writes are committed and are what
travels to the GPU box.
"""
import io
import json
import os
import sys
import zlib
from contextlib import redirect_stdout

import numpy as np
import scipy.io as sio
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

from refpykrylov.linop import LinearOperator, DiagonalOperator      # noqa: E402
from refpykrylov.cg import CG                                       # noqa: E402
from refpykrylov.cgs import CGS                                     # noqa: E402
from refpykrylov.tfqmr import TFQMR                                 # noqa: E402
from refpykrylov.bicgstab import BiCGSTAB                           # noqa: E402
from refpykrylov.minres import Minres                               # noqa: E402
from refpykrylov.gallery import Poisson1dMatvec, Poisson2dMatvec    # noqa: E402

K_HIST = 25


def csr(path):
    M = sp.csr_matrix(sio.mmread(os.path.join(HERE, path)))
    M.sort_indices()
    return M


def op_of(M, symmetric=False):
    return LinearOperator(M.shape[1], M.shape[0], lambda v: M @ v,
                          matvec_transp=lambda u: M.T @ u, symmetric=symmetric)


def record(k, x=None, e=None):
    out = dict(nMatvec=int(k.nMatvec), residNorm0=float(k.residNorm0), residNorm=float(k.residNorm),
               converged=bool(k.converged), residHistory=[float(v) for v in k.residHistory[:K_HIST]])
    if e is not None:
        out["err"] = float(np.linalg.norm(k.bestSolution - e) / np.sqrt(len(e)))
    return out


def main():
    G, V = {}, {}
    # ---- integer work: CSR of the fixtures as scipy builds them
    for name in ("1138bus", "jpwh_991", "GD97_b"):
        M = csr(name + ".mtx")
        G["csr/" + name] = dict(shape=list(M.shape), nnz=int(M.nnz),
                                indptr_crc=zlib.crc32(M.indptr.astype(np.int32).tobytes()),
                                indices_crc=zlib.crc32(M.indices.astype(np.int32).tobytes()),
                                data_crc=zlib.crc32(M.data.astype(np.float64).tobytes()))
    # ---- CG known answers (pykrylov/cg/tests/test_diagdom.py, doc/source/introduction.rst:46-48)
    for n in (10, 20, 100, 1000):
        A = LinearOperator(n, n, lambda x: Poisson1dMatvec(x), symmetric=True)
        e = np.ones(n)
        cg = CG(A, matvec_max=2 * n)
        cg.solve(A * e)
        G["cg/poisson1d/%d" % n] = record(cg, e=e)
    for g in (10, 20, 100):
        n2 = g * g
        A = LinearOperator(n2, n2, lambda x: Poisson2dMatvec(x), symmetric=True)
        e = np.ones(n2)
        cg = CG(A, matvec_max=2 * n2)
        cg.solve(A * e)
        G["cg/poisson2d_gallery/%d" % g] = record(cg, e=e)
    # the same operators as CSR (what the device iterates on)
    for g in (10, 20, 100):
        n2 = g * g
        T = sp.diags([-1, 2, -1], [-1, 0, 1], shape=(g, g))
        M = (sp.kron(sp.identity(g), T) + sp.kron(T, sp.identity(g))).tocsr()
        M.sort_indices()
        A = op_of(M, symmetric=True)
        e = np.ones(n2)
        cg = CG(A)
        cg.solve(M @ e)
        G["cg/poisson2d_csr/%d" % g] = record(cg, e=e)
        if g == 20:
            V["cg_poisson2d_csr_20_x"] = cg.bestSolution.copy()
    # ---- CG on 1138bus (doc/source/cg.rst:59), default tolerances, guess 1..n like demo_common.py
    M = csr("1138bus.mtx")
    n = M.shape[0]
    e = np.ones(n)
    cg = CG(op_of(M, True), reltol=1.0e-8)
    cg.solve(M @ e, guess=1 + np.arange(n, dtype=float), matvec_max=2 * n)
    G["cg/1138bus/demo"] = record(cg, e=e)
    cg = CG(op_of(M, True))
    cg.solve(M @ e)
    G["cg/1138bus/default"] = record(cg, e=e)
    # diagonally preconditioned CG (DiagonalOperator precon)
    dinv = 1.0 / M.diagonal()
    cg = CG(op_of(M, True), precon=DiagonalOperator(dinv))
    cg.solve(M @ e)
    G["cg/1138bus/jacobi"] = record(cg, e=e)
    # ---- bmark.py (doc/source/bmark.rst:52-54): jpwh_991, reltol 1e-8, guess 1..n, matvec_max 2n
    M = csr("jpwh_991.mtx")
    n = M.shape[0]
    e = np.ones(n)
    rhs = M @ e
    for K in (CGS, TFQMR, BiCGSTAB):
        for reltol in (1.0e-8, 1.0e-5):
            ks = K(op_of(M), reltol=reltol)
            ks.solve(rhs, guess=1 + np.arange(n, dtype=float), matvec_max=2 * n)
            key = "%s/jpwh_991/reltol%g" % (ks.acronym, reltol)
            G[key] = record(ks, e=e)
            V[key.replace("/", "_") + "_x"] = ks.bestSolution.copy()
    # zero initial guess, random rhs (rhs = A*ones is an exact Bi-CGSTAB breakdown here)
    b = M @ np.random.default_rng(5).standard_normal(n)
    for K in (CGS, TFQMR, BiCGSTAB):
        ks = K(op_of(M), reltol=1.0e-8)
        ks.solve(b, matvec_max=2 * n)
        G["%s/jpwh_991/zero_guess" % ks.acronym] = record(ks)
    # ---- MINRES on S = (B + B^T)/2, B = jpwh_991 (one block)
    S = ((M + M.T) * 0.5).tocsr()
    S.sort_indices()
    mr = Minres(op_of(S, True))
    with redirect_stdout(io.StringIO()):
        mr.solve(S @ e, show=False)
    G["MINRES/sym_jpwh_991"] = dict(istop=int(mr.istop), itn=int(mr.itn), rnorm=float(mr.rnorm),
                                    Anorm=float(mr.Anorm), Acond=float(mr.Acond),
                                    ynorm=float(mr.ynorm), Arnorm=float(mr.Arnorm),
                                    residNorm0=float(mr.residNorm0),
                                    residHistory=[float(v) for v in mr.residHistory[:K_HIST]],
                                    dir_errors_window=[float(v) for v in mr.dir_errors_window[:K_HIST]],
                                    err=float(np.linalg.norm(mr.x - e) / np.sqrt(n)))
    V["MINRES_sym_jpwh_991_x"] = mr.x.copy()
    # nonsymmetric operator: symmetry check must stop MINRES at 0 iterations with istop 7
    mr = Minres(op_of(M))
    with redirect_stdout(io.StringIO()):
        mr.solve(rhs, show=False)
    G["MINRES/jpwh_991_nonsym"] = dict(istop=int(mr.istop), itn=int(mr.itn))
    # shifted system
    mr = Minres(op_of(S, True))
    with redirect_stdout(io.StringIO()):
        mr.solve(S @ e, show=False, shift=0.5, check=False)
    G["MINRES/sym_jpwh_991_shift0.5"] = dict(istop=int(mr.istop), itn=int(mr.itn), rnorm=float(mr.rnorm),
                                             residHistory=[float(v) for v in mr.residHistory[:K_HIST]])

    with open(os.path.join(HERE, "golden.json"), "w") as fh:
        json.dump(G, fh, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "golden_vectors.npz"), **V)
    print("wrote %d records, %d vectors" % (len(G), len(V)))
    for k in sorted(G):
        v = G[k]
        print("%-40s %s" % (k, {kk: vv for kk, vv in v.items() if kk not in ("residHistory", "dir_errors_window")}))


if __name__ == "__main__":
    main()
