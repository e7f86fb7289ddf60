#!/usr/bin/env python3
"""Generate tests/golden/golden_lls.json + golden_lls_vectors.npz FROM THE REFERENCE ITSELF.

Runs the transliterated reference (oracle/_ref/refpykrylov, produced by oracle/make_ref.py from
/root/reference) -- LSQR, LSMR, CRAIG, CRAIG-MR (pykrylov/lls) and SYMMLQ (pykrylov/symmlq) --
through its own public API with scipy-CSR stand-in operators, and records the scalar outcomes,
the first K_HIST entries of every per-iteration history the reference keeps, and the final
iterates.  Only runs in the build container; the files it writes are committed and are what
travels to the GPU box (tests/test_gpu_lls.py, tests/test_host_solvers.py).
"""
import io
import json
import os
import sys
from contextlib import redirect_stdout

import numpy as np
import scipy.io as sio
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))

from refpykrylov.linop import LinearOperator                          # noqa: E402
from refpykrylov.lls import LSQRFramework, LSMRFramework, CRAIGFramework, CRAIGMRFramework   # noqa: E402
from refpykrylov.symmlq import Symmlq                                 # noqa: E402

K_HIST = 25


def csr(name):
    M = sp.csr_matrix(sio.mmread(os.path.join(HERE, name + ".mtx")))
    M.sort_indices()
    return M


def op_of(M, symmetric=False):
    return LinearOperator(M.shape[1], M.shape[0], lambda v: M @ v, matvec_transp=lambda u: M.T @ u,
                          symmetric=symmetric)


def fl(seq):
    return [float(v) for v in list(seq)[:K_HIST]]


def quiet(fun, *a, **kw):
    with redirect_stdout(io.StringIO()):
        return fun(*a, **kw)


def main():
    G, V = {}, {}
    # ---- LSQR on jpwh_991 (square, nonsymmetric), with and without damping
    J = csr("jpwh_991")
    n = J.shape[0]
    rhs = J @ np.ones(n)
    for damp in (0.0, 0.1):
        ls = LSQRFramework(op_of(J))
        quiet(ls.solve, rhs, damp=damp, store_resids=True)
        key = "LSQR/jpwh_991/damp%g" % damp
        G[key] = dict(istop=int(ls.istop), itn=int(ls.itn), r1norm=float(ls.r1norm), r2norm=float(ls.r2norm),
                      Anorm=float(ls.Anorm), Acond=float(ls.Acond), Arnorm=float(ls.Arnorm), xnorm=float(ls.xnorm),
                      err=float(np.linalg.norm(ls.x - 1.0) / np.sqrt(n)), resids=fl(ls.resids),
                      normal_eqns_resids=fl(ls.normal_eqns_resids), dir_errors_window=fl(ls.dir_errors_window))
        V[key.replace("/", "_") + "_x"] = ls.x
    # ---- over-determined random least squares: LSQR and LSMR
    R = sp.random(600, 200, density=0.03, random_state=7, format="csr")
    R.sort_indices()
    b = np.random.default_rng(7).standard_normal(600)
    ls = LSQRFramework(op_of(R))
    quiet(ls.solve, b, store_resids=True)
    G["LSQR/random_600x200"] = dict(istop=int(ls.istop), itn=int(ls.itn), r1norm=float(ls.r1norm),
                                    r2norm=float(ls.r2norm), Anorm=float(ls.Anorm), Acond=float(ls.Acond),
                                    Arnorm=float(ls.Arnorm), xnorm=float(ls.xnorm), resids=fl(ls.resids),
                                    normal_eqns_resids=fl(ls.normal_eqns_resids),
                                    dir_errors_window=fl(ls.dir_errors_window))
    V["LSQR_random_600x200_x"] = ls.x
    lm = LSMRFramework(op_of(R))
    x, istop, itn, normr, normar, normA, condA, normx = quiet(lm.solve, b, store_resids=True)
    G["LSMR/random_600x200"] = dict(istop=int(istop), itn=int(itn), normr=float(normr), normar=float(normar),
                                    normA=float(normA), condA=float(condA), normx=float(normx),
                                    resids=fl(lm.resids), normal_eqns_resids=fl(lm.normal_eqns_resids),
                                    norms=fl(lm.norms), dir_errors_window=fl(lm.dir_errors_window))
    V["LSMR_random_600x200_x"] = x
    # ---- consistent under-determined system: CRAIG and CRAIG-MR (least-norm solvers)
    rng = np.random.default_rng(14)
    U = sp.random(90, 150, density=0.1, random_state=5, format="csr")
    U.sort_indices()
    bu = U @ rng.standard_normal(150)
    cr = CRAIGFramework(op_of(U))
    quiet(cr.solve, bu, store_resids=True)
    G["CRAIG/random_90x150"] = dict(istop=int(cr.istop), itn=int(cr.itn), r1norm=float(cr.r1norm),
                                    r2norm=float(cr.r2norm), Arnorm=float(cr.Arnorm), xnorm=float(cr.xnorm),
                                    resids=fl(cr.resids), normal_eqns_resids=fl(cr.normal_eqns_resids),
                                    norms=fl(cr.norms), dir_errors_d_window=fl(cr.dir_errors_d_window))
    V["CRAIG_random_90x150_x"] = cr.x
    cm = CRAIGMRFramework(op_of(U))
    quiet(cm.solve, bu, store_resids=True)
    G["CRAIGMR/random_90x150"] = dict(istop=int(cm.istop), itn=int(cm.itn), norms=fl(cm.norms),
                                      normal_eqns_resids=fl(cm.normal_eqns_resids),
                                      dir_errors_window=fl(cm.dir_errors_window))
    V["CRAIGMR_random_90x150_x"] = cm.x
    # ---- SYMMLQ on the symmetric part of jpwh_991 (with and without shift) and on 1138bus (capped)
    S = ((J + J.T) * 0.5).tocsr()
    S.sort_indices()
    rs = S @ np.ones(n)
    for shift, key in ((None, "SYMMLQ/sym_jpwh_991"), (0.5, "SYMMLQ/sym_jpwh_991_shift0.5")):
        sq = Symmlq(op_of(S, symmetric=True))
        quiet(sq.solve, rs, **({} if shift is None else {"shift": shift}))
        G[key] = dict(nMatvec=int(sq.nMatvec), residNorm=float(sq.residNorm), xNorm=float(sq.xNorm),
                      anorm=float(sq.anorm), acond=float(sq.acond))
        V[key.replace("/", "_") + "_x"] = sq.x
    Bm = csr("1138bus")
    sq = Symmlq(op_of(Bm, symmetric=True))
    quiet(sq.solve, Bm @ np.ones(Bm.shape[0]), matvec_max=400)
    G["SYMMLQ/1138bus_max400"] = dict(nMatvec=int(sq.nMatvec), residNorm=float(sq.residNorm), xNorm=float(sq.xNorm),
                                      anorm=float(sq.anorm), acond=float(sq.acond))
    with open(os.path.join(HERE, "golden_lls.json"), "w") as fh:
        json.dump(G, fh, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "golden_lls_vectors.npz"), **V)
    print("wrote %d records, %d vectors" % (len(G), len(V)))


if __name__ == "__main__":
    main()
