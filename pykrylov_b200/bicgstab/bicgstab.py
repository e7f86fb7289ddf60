"""Bi-CGSTAB, device-resident (reference: pykrylov/bicgstab/bicgstab.py:7-151).

The loop of bicgstab.py:85-145 runs on the GPU as 4 fused launches per iteration
(2 SpMV+dot kernels, 2 multi-AXPY+norm kernels); the four stopping tests of the
reference are evaluated on device.
"""
import numpy as np

from ..generic import KrylovMethod
from .. import _engine

__docformat__ = "restructuredtext"


class BiCGSTAB(KrylovMethod):
    """Bi-conjugate gradient stabilised for unsymmetric nonsingular ``A x = b``
    (2 operator products, 6 inner products, 6 AXPYs per iteration; never uses A^T)."""

    def __init__(self, op, **kwargs):
        KrylovMethod.__init__(self, op, **kwargs)
        self.name = "Bi-Conjugate Gradient Stabilized"
        self.acronym = "Bi-CGSTAB"
        self.prefix = self.acronym + ": "

    def solve(self, rhs, **kwargs):
        """Keywords (bicgstab.py:47-50): ``guess`` (default 0), ``matvec_max`` (2n)."""
        n = rhs.shape[0]
        result_type = _engine.check_real(self.op, rhs)
        guess = kwargs.get("guess", None)
        matvec_max = kwargs.get("matvec_max", 2 * _engine.global_size(self.op, n))
        plan = _engine.resolve(self.op, self.precon, n)
        if plan is None:       # closure operator / opaque preconditioner: host-driven loop on device vectors
            from .. import _bridged
            return _bridged.bicgstab(self, rhs, guess, matvec_max, result_type)
        S = _engine.make_solver("bicgstab", plan, self.context)
        S.setup(rhs, guess=guess, abstol=self.abstol, reltol=self.reltol, matvec_max=matvec_max)
        state = {"first": True, "nmv": 1 if guess is not None else 0}

        def replay(st, hist):
            for (resid,) in hist:
                if state["first"]:
                    state["first"] = False
                    self.residNorm0 = resid
                    self.logger.info("Initial residual = %8.2e" % resid)
                    self.logger.info("Threshold = %8.2e" % st.threshold)
                    hdr = "%6s  %8s" % ("Matvec", "Residual")
                    self.logger.info(hdr)
                    self.logger.info("-" * len(hdr))
                else:
                    state["nmv"] += 1
                    self.logger.info("%6d  %8.2e" % (state["nmv"], resid))

        st = _engine.drive(S, self.check_interval, replay)
        self.converged = bool(st.resid_norm <= st.threshold)
        self.nMatvec = int(st.n_matvec)
        self.bestSolution = self.x = S.solution().astype(result_type, copy=False)
        self.residNorm = st.resid_norm
        self.op._nMatvec += self.nMatvec
