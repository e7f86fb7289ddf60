"""MINRES, device-resident (reference: pykrylov/minres/minres.py:14-408).

Keyword contract of the reference's ``solve`` (minres.py:121-130): ``precon``,
``shift``, ``show`` (True), ``check`` (True: 20 operator products for the
randomised symmetry test), ``itnlim`` (5n), ``rtol`` (1e-12), ``etol`` (1e-6),
``store_resids``, ``store_iterates``, ``window`` (5).  The Lanczos step
(minres.py:236-256), the QR update (:270-297) and all stopping tests (:323-361)
run on the GPU, 3 fused launches per iteration; so does a diagonal preconditioner
(``DiagonalOperator`` or a bmark-style ``r / diag`` object).  Closure operators and opaque
preconditioners go through the host-callback bridge.
"""
import numpy as np

from ..generic import KrylovMethod
from ..tools.utils import check_symmetric, machine_epsilon
from .. import _engine

__docformat__ = "restructuredtext"


class Minres(KrylovMethod):
    """Paige-Saunders MINRES for symmetric (possibly indefinite / singular)
    ``(A - shift I) x = b``."""

    msg = [" beta2 = 0.  If M = I, b and x are eigenvectors    ",     # -1
           " beta1 = 0.  The exact solution is  x = 0          ",     # 0
           " A solution to Ax = b was found, given rtol        ",     # 1
           " A least-squares solution was found, given rtol    ",     # 2
           " Reasonable accuracy achieved, given eps           ",     # 3
           " x has converged to an eigenvector                 ",     # 4
           " acond has exceeded 0.1/eps                        ",     # 5
           " The iteration limit was reached                   ",     # 6
           " Aname  does not define a symmetric matrix         ",     # 7
           " Mname  does not define a symmetric matrix         ",     # 8
           " Mname  does not define a pos-def preconditioner   ",     # 9
           "The truncated direct error is small enough, given etol"]  # 10

    def __init__(self, op, **kwargs):
        self.first = "Enter minres.   "
        self.last = "Exit  minres.   "
        KrylovMethod.__init__(self, op, **kwargs)
        self.name = "Minimum Residual"
        self.acronym = "MINRES"
        self.prefix = self.acronym + ": "
        self.residHistory = []
        self.resids = []
        self.dir_errors_window = []
        self.iterates = []
        self.eps = machine_epsilon()

    def normof2(self, x, y):
        return np.sqrt(x ** 2 + y ** 2)

    def solve(self, b, **kwargs):
        A = self.op
        n = b.shape[0]
        precon = kwargs.get("precon", None)
        shift = kwargs.get("shift", 0.0)
        show = kwargs.get("show", True)
        check = kwargs.get("check", True)
        itnlim = kwargs.get("itnlim", 5 * _engine.global_size(A, n))
        rtol = kwargs.get("rtol", 1.0e-12)
        etol = kwargs.get("etol", 1.0e-6)
        store_iterates = kwargs.get("store_iterates", False)
        window = kwargs.get("window", 5)
        self.dir_errors_window = []
        self.iterates = []
        result_type = _engine.check_real(A, b)
        plan = _engine.resolve(A, precon, n)        # device operator + no / a diagonal preconditioner

        if show:
            print(self.first + "Solution of symmetric Ax = b")
            print("n      =  %3d     precon =  %4s           shift  =  %23.14e"
                  % (n, (precon is not None), shift))
            print("itnlim =  %3d     rtol   =  %11.2e\n" % (itnlim, rtol))

        if plan is None:       # closure operator or an opaque preconditioner: host-driven loop on device vectors
            from .. import _bridged
            return _bridged.minres(self, b, precon, shift, show, check, itnlim, rtol, etol,
                                   store_iterates, window, result_type)
        S = _engine.make_solver("minres", plan, self.context)
        symmetric_ok = True
        if check:                                         # minres.py:186-189
            symmetric_ok = check_symmetric(A)
        S.setup(b, abstol=0.0, reltol=0.0, matvec_max=itnlim if symmetric_ok else 0,
                shift=shift, rtol=rtol, etol=etol, window=window)
        if show:
            print(" " * 2)
            print("   Itn     x[0]     Compatible    LS" + "       norm(A)  cond(A) gbar/|A|")

        def replay(st, hist):
            for rnorm, direrr in hist:
                self.residHistory.append(rnorm)
                if direrr == direrr:                      # NaN until itn > window (minres.py:305)
                    self.dir_errors_window.append(direrr)
            if store_iterates:
                self.iterates.append(S.solution())

        interval = 1 if store_iterates else self.check_interval
        st = _engine.drive(S, interval, replay, overlap=not store_iterates)
        istop = int(st.istop)
        if not symmetric_ok:
            istop = 7
        itn = int(st.n_iter)
        Anorm, Acond, ynorm, Arnorm = st.aux[0], st.aux[1], st.aux[2], st.aux[3]
        rnorm = st.resid_norm
        if show:
            last = self.last
            print(last + " istop   =  %3g               itn   =%5g" % (istop, itn))
            print(last + " Anorm   =  %12.4e      Acond =  %12.4e" % (Anorm, Acond))
            print(last + " rnorm   =  %12.4e      ynorm =  %12.4e" % (rnorm, ynorm))
            print(last + " Arnorm  =  %12.4e" % Arnorm)
            print(last + self.msg[istop + 1])
        self.residNorm0 = st.resid_norm0
        self.converged = istop in [1, 2, 3, 4, 10]
        if istop == 10:
            self.status = "direct error small"
        self.x = self.bestSolution = S.solution().astype(result_type, copy=False)
        self.istop = istop
        self.itn = self.nMatvec = itn
        self.rnorm = self.residNorm = rnorm
        self.Arnorm = Arnorm
        self.Anorm = Anorm
        self.Acond = Acond
        self.ynorm = ynorm
        self.op._nMatvec += itn
        return
