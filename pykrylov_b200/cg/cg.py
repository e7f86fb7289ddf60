"""Conjugate gradient, device-resident.

Same class surface and keyword contract as the reference's pykrylov/cg/cg.py:7-165;
the loop of cg.py:113-158 itself runs on the GPU (libkrylov_b200, 2 fused launches
per iteration; 3 on row shards) with the stopping test evaluated on device every iteration.  The
host only replays the log lines / ``residHistory`` from the device ring buffer at
every ``check_interval``.
"""
import numpy as np

from ..generic import KrylovMethod
from ..tools.utils import check_symmetric
from .. import _engine

__docformat__ = "restructuredtext"


class CG(KrylovMethod):
    """Conjugate gradient for symmetric positive definite ``A x = b``
    (1 operator product, 2 inner products, 3 AXPYs per iteration)."""

    def __init__(self, op, **kwargs):
        KrylovMethod.__init__(self, op, **kwargs)
        self.name = "Conjugate Gradient"
        self.acronym = "CG"
        self.prefix = self.acronym + ": "
        self.resids = []
        self.iterates = []
        self.infiniteDescent = None

    def solve(self, rhs, **kwargs):
        """Solve ``A x = rhs``.  Keywords (cg.py:57-67): ``guess``, ``matvec_max``
        (default 2n), ``check_symmetric`` (False), ``check_curvature`` (True),
        ``store_resids`` (False), ``store_iterates`` (False)."""
        n = rhs.shape[0]
        check_sym = kwargs.get("check_symmetric", False)
        check_curvature = kwargs.get("check_curvature", True)
        store_resids = kwargs.get("store_resids", False)
        store_iterates = kwargs.get("store_iterates", False)

        if check_sym and not check_symmetric(self.op):
            self.logger.error("Coefficient operator is not symmetric")
            return

        result_type = _engine.check_real(self.op, rhs)
        guess = kwargs.get("guess", None)
        matvec_max = kwargs.get("matvec_max", 2 * _engine.global_size(self.op, n))
        plan = _engine.resolve(self.op, self.precon, n)
        if plan is None:
            return self._solve_bridged(rhs, guess, matvec_max, check_curvature,
                                       store_resids, store_iterates, result_type)

        S = _engine.make_solver("cg", plan, self.context)
        S.setup(rhs, guess=guess, abstol=self.abstol, reltol=self.reltol, matvec_max=matvec_max,
                check_curvature=check_curvature)
        interval = 1 if (store_resids or store_iterates) else self.check_interval
        state = {"first": True, "nmv": 0}
        hdr = "%6s  %7s  %8s" % ("Matvec", "Resid", "Curv")

        def replay(st, hist):
            for resid, curv in hist:
                if state["first"]:
                    state["first"] = False
                    state["nmv"] = 1 if guess is not None else 0
                    self.residNorm0 = resid
                    self.logger.info(hdr)
                    self.logger.info("-" * len(hdr))
                    self.logger.info("%6d  %7.1e" % (state["nmv"], resid))
                else:
                    state["nmv"] += 1
                    self.logger.info("%6d  %7.1e  %8.1e" % (state["nmv"], resid, curv))
                self.residHistory.append(resid)
            if store_iterates:
                self.iterates.append(S.solution())
            if store_resids:
                r = S.get_vector("r")
                self.resids.append(r if plan.precon_mode == 0 else
                                   (plan.precon_diag * r if plan.precon_mode == 1 else r / plan.precon_diag))

        st = _engine.drive(S, interval, replay, overlap=not (store_resids or store_iterates))
        if not st.definite:
            self.logger.error("Coefficient operator is not positive definite")
            self.infiniteDescent = S.get_vector("p")
        self.converged = bool(st.resid_norm <= st.threshold)
        self.definite = bool(st.definite)
        self.nMatvec = int(st.n_matvec)
        self.bestSolution = self.x = S.solution().astype(result_type, copy=False)
        self.residNorm = st.resid_norm
        self.op._nMatvec += self.nMatvec

    # -- opaque operator / preconditioner: vectors in HBM, operator through the host
    def _solve_bridged(self, rhs, guess, matvec_max, check_curvature, store_resids,
                       store_iterates, result_type):
        n = rhs.shape[0]
        B = _engine.HostBridge(n, self.context, self.op)
        op, precon = self.op, self.precon
        nMatvec = 0
        definite = True
        x = B.vec(guess)
        if store_iterates:
            self.iterates.append(x.download())
        r = B.vec(-np.asarray(rhs, dtype=np.float64))
        Ap = B.vec()
        if guess is not None:
            B.apply(op, x, Ap)
            nMatvec += 1
            B.fused([dict(z=r, u=r, w=Ap, a=1.0, b=1.0)])                        # r += A x
        y = r if precon is None else B.apply_precon(precon, r, B.vec())
        if store_resids:
            self.resids.append(y.download())
        ry = B.fused([], [(r, y)])[0]
        self.residNorm0 = residNorm = np.abs(np.sqrt(ry))
        self.residHistory.append(self.residNorm0)
        threshold = max(self.abstol, self.reltol * self.residNorm0)
        p = B.vec()
        B.fused([dict(z=p, u=r, a=-1.0)])                                        # p = -r
        hdr = "%6s  %7s  %8s" % ("Matvec", "Resid", "Curv")
        self.logger.info(hdr)
        self.logger.info("-" * len(hdr))
        self.logger.info("%6d  %7.1e" % (nMatvec, residNorm))
        while residNorm > threshold and nMatvec < matvec_max and definite:
            B.apply(op, p, Ap)
            nMatvec += 1
            pAp = B.fused([], [(p, Ap)])[0]
            if check_curvature and pAp <= 0:
                self.logger.error("Coefficient operator is not positive definite")
                self.infiniteDescent = p.download()
                definite = False
                continue
            alpha = ry / pAp
            ops = [dict(z=x, u=x, w=p, a=1.0, b=alpha), dict(z=r, u=r, w=Ap, a=1.0, b=alpha)]
            if precon is None:
                ry_next = B.fused(ops, [(r, r)])[0]
            else:
                B.fused(ops)
                B.apply_precon(precon, r, y)
                ry_next = B.fused([], [(r, y)])[0]
            if store_iterates:
                self.iterates.append(x.download())
            if store_resids:
                self.resids.append(y.download())
            beta = ry_next / ry
            B.fused([dict(z=p, u=p, w=r, a=beta, b=-1.0)])                       # p = beta p - r
            ry = ry_next
            residNorm = np.abs(np.sqrt(ry))
            self.residHistory.append(residNorm)
            self.logger.info("%6d  %7.1e  %8.1e" % (nMatvec, residNorm, pAp))
        self.converged = bool(residNorm <= threshold)
        self.definite = definite
        self.nMatvec = nMatvec
        self.bestSolution = self.x = x.download().astype(result_type, copy=False)
        self.residNorm = residNorm
