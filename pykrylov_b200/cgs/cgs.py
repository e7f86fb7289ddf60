"""Conjugate gradient squared, device-resident (reference: pykrylov/cgs/cgs.py:7-123).

The loop of cgs.py:76-117 runs on the GPU as 4 fused launches per iteration.
Like the reference, the product that forms the initial residual of a supplied
guess is *not* counted in ``nMatvec`` (cgs.py:59-60).
"""
from ..generic import KrylovMethod
from .. import _engine

__docformat__ = "restructuredtext"


class CGS(KrylovMethod):
    """CGS for unsymmetric ``A x = b`` (2 operator products, 3 inner products,
    7 AXPYs per iteration; never uses A^T)."""

    def __init__(self, op, **kwargs):
        KrylovMethod.__init__(self, op, **kwargs)
        self.name = "Conjugate Gradient Squared"
        self.acronym = "CGS"
        self.prefix = self.acronym + ": "

    def solve(self, rhs, **kwargs):
        """Keywords (cgs.py:45-48): ``guess`` (default 0), ``matvec_max`` (2n)."""
        n = rhs.shape[0]
        result_type = _engine.check_real(self.op, rhs)
        guess = kwargs.get("guess", None)
        matvec_max = kwargs.get("matvec_max", 2 * _engine.global_size(self.op, n))
        plan = _engine.resolve(self.op, self.precon, n)
        if plan is None:       # closure operator / opaque preconditioner: host-driven loop on device vectors
            from .. import _bridged
            return _bridged.cgs(self, rhs, guess, matvec_max, result_type)
        S = _engine.make_solver("cgs", plan, self.context)
        S.setup(rhs, guess=guess, abstol=self.abstol, reltol=self.reltol, matvec_max=matvec_max)
        state = {"first": True, "nmv": 0}

        def replay(st, hist):
            for (resid,) in hist:
                if state["first"]:
                    state["first"] = False
                    self.residNorm0 = resid
                    self.logger.info("Initial residual = %8.2e\n" % resid)
                    self.logger.info("Threshold = %8.2e\n" % st.threshold)
                else:
                    state["nmv"] += 2
                    state["pending"] = "%5d  %8.2e\n" % (state["nmv"], resid)
                    # the reference logs only iterations that continue (cgs.py:102-117)
                    if not (resid <= st.threshold or state["nmv"] >= matvec_max):
                        self.logger.info(state["pending"])

        st = _engine.drive(S, self.check_interval, replay)
        self.converged = bool(st.resid_norm <= st.threshold)
        self.nMatvec = int(st.n_matvec)
        self.bestSolution = self.x = S.solution().astype(result_type, copy=False)
        self.residNorm = st.resid_norm
        self.op._nMatvec += self.nMatvec + (1 if guess is not None else 0)
