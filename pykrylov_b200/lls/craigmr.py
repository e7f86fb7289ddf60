"""CRAIG-MR on device vectors (reference: pykrylov/lls/craigmr.py:13-250).

Minimum-residual variant of the generalised CRAIG method; the iterate ``x`` it returns is
the *dual* variable of size m.  Quirks kept from the reference: every iteration prints
``itn xNrgNorm2`` unconditionally (craigmr.py:190) and no residual-based stopping test
exists (only the truncated direct error and the iteration limit).  Vector work (the
Golub-Kahan step with A and the device-built A^T, the ``d / dbar / x`` recurrences,
craigmr.py:147-193) runs as CUDA kernels; the rotations stay on the host.
"""
from math import sqrt

import numpy as np

from ..generic import KrylovMethod
from .. import _engine

__docformat__ = "restructuredtext"


class CRAIGMRFramework(KrylovMethod):

    msg = ("The exact solution is  x = 0                              ",
           "Ax - b is small enough, given atol, btol                  ",
           "The least-squares solution is good enough, given atol     ",
           "The estimate of cond(Abar) has exceeded conlim            ",
           "Ax - b is small enough for this machine                   ",
           "The least-squares solution is good enough for this machine",
           "Cond(Abar) seems to be too large for this machine         ",
           "The iteration limit has been reached                      ",
           "The truncated direct error is small enough, given etol    ")

    def __init__(self, A, **kwargs):
        KrylovMethod.__init__(self, A, **kwargs)
        self.name = "Least-Norm Minimum Residual"
        self.acronym = "CRAIG-MR"
        self.prefix = self.acronym + ": "
        self.A = A
        self.init_data()

    def init_data(self):
        self.x = None
        self.var = None
        self.itn = 0
        self.istop = 0
        self.Anorm = self.Acond = self.Arnorm = self.xnorm = 0.
        self.r1norm = self.r2norm = 0.
        self.optimal = False
        self.resids = []
        self.normal_eqns_resids = []
        self.norms = []
        self.dir_errors_window = []
        self.iterates = []

    def solve(self, b, damp=0.0, atol=1e-9, btol=1e-9, conlim=1e8, M=None, N=None, itnlim=None,
              show=False, **kwargs):
        etol = kwargs.get("etol", 1.0e-6)
        store_resids = kwargs.get("store_resids", False)
        store_iterates = kwargs.get("store_iterates", False)
        window = kwargs.get("window", 5)
        self.init_data()
        A = self.A
        b = np.asarray(b, dtype=np.float64).squeeze()
        m, n = A.shape
        minDim = min([m, n])
        if itnlim is None:
            itnlim = minDim
        B = _engine.HostBridge(n, self.context, A)
        Mu = B.vec_n(m, b)
        u = Mu if M is None else B.apply_callable(M, Mu, B.vec_n(m))
        beta = sqrt(B.fused([], [(u, Mu)])[0])
        Nv = B.vec_n(n)
        v = Nv if N is None else B.vec_n(n)
        tm, tn = B.vec_n(m), B.vec_n(n)
        alpha = 0
        if beta > 0:
            ops = [dict(z=u, u=u, a=beta, a_div=True)]
            if M is not None:
                ops.append(dict(z=Mu, u=Mu, a=beta, a_div=True))
            B.fused(ops)
            B.apply(A, u, Nv, trans=True)
            if N is not None:
                B.apply_callable(N, Nv, v)
            alpha = sqrt(B.fused([], [(v, Nv)])[0])
        if alpha > 0:
            ops = [dict(z=v, u=v, a=alpha, a_div=True)]
            if N is not None:
                ops.append(dict(z=Nv, u=Nv, a=alpha, a_div=True))
            B.fused(ops)
        itn = 0
        alpha_hat = sqrt(alpha ** 2 + 1)
        c = alpha / alpha_hat
        s = 1. / alpha_hat
        zeta_hat = beta
        alpha_tilde = alpha_hat
        theta = 0.
        d, dbar, x = B.vec_n(m), B.vec_n(m), B.vec_n(m)
        B.fused([dict(z=d, u=u, a=alpha_hat, a_div=True)])
        if store_iterates:
            self.iterates.append(x.download())
        xNrgNorm2 = 0.
        dErr = np.zeros(window)
        trncDirErr = 0
        istop = 0
        if store_resids:
            self.norms.append(xNrgNorm2)
        csr = _engine.plane_csr(A)
        on_device = (csr is not None and M is None and N is None and not store_iterates and itnlim > 0)
        if on_device:
            # device-resident loop (scalar plane in csrc/lls.cu, see lsqr.py)
            from ..device import ScalarPlane as SL
            loop = _engine.PlaneLoop(B.ctx, "craigmr")
            loop.P.setup(dict(alpha=alpha, beta=beta, c=c, s=s, zeta_hat=zeta_hat, alpha_tilde=alpha_tilde, theta=theta,
                              xNrgNorm2=xNrgNorm2), window=window, itnlim=itnlim, etol=etol)

            def trip():
                # Mu = A v - alpha Mu, |Mu|^2 and phase 1 (beta, |A|) in one launch; likewise A'u below
                loop.spmv_ops(csr, v, dict(z=Mu, w=Mu, a=1.0, b_slot=SL.ALPHA, b_neg=1), step=1)
                loop.ops([dict(z=u, u=u, a_slot=SL.U_DIV, a_div=True)])
                loop.spmv_ops(csr, u, dict(z=Nv, w=Nv, a_slot=SL.NV_A, b_slot=SL.NV_B), step=2, trans=True)
                loop.ops([dict(z=v, u=v, a_slot=SL.V_DIV, a_div=True)])
                loop.ops([dict(z=dbar, u=d, w=dbar, a=1.0, b_slot=SL.C0, b_neg=1), dict(z=dbar, u=dbar, a_slot=SL.C1, a_div=True),
                          dict(z=d, u=u, w=d, a=1.0, b_slot=SL.C2, b_neg=1), dict(z=d, u=d, a_slot=SL.C3, a_div=True)])
                loop.ops([dict(z=x, u=x, w=dbar, a=1.0, b_slot=SL.C4)])
                loop.P.step(9)                              # latch `done` behind this trip's updates

            seen = [0]

            def replay(st_, sc_, hist):
                for nrg, azeta, direrr, _ in hist:
                    seen[0] += 1
                    print(seen[0], nrg)                                     # (sic) craigmr.py:190
                    if store_resids:
                        self.norms.append(nrg)
                        self.normal_eqns_resids.append(azeta)
                    if direrr == direrr:
                        self.dir_errors_window.append(direrr)

            if not csr.symmetric:
                csr.build_transpose()
            st_, sc = loop.run(trip, self.check_interval, replay)
            itn, istop = int(st_.itn), int(st_.istop)
            xNrgNorm2, trncDirErr = sc["xNrgNorm2"], sc["trncDirErr"]
            A._nMatvec += 2 * itn

        while itn < itnlim and not on_device:
            itn += 1
            B.apply(A, v, tm)
            if M is None:
                beta = sqrt(B.fused([dict(z=Mu, u=tm, w=Mu, a=1.0, b=-alpha)], [(Mu, Mu)])[0])
            else:
                B.fused([dict(z=Mu, u=tm, w=Mu, a=1.0, b=-alpha)])
                B.apply_callable(M, Mu, u)
                beta = sqrt(B.fused([], [(u, Mu)])[0])
            if beta > 0:
                ops = [dict(z=u, u=u, a=beta, a_div=True)]
                if M is not None:
                    ops.append(dict(z=Mu, u=Mu, a=beta, a_div=True))
                B.fused(ops)
                B.apply(A, u, tn, trans=True)
                if N is None:
                    alpha = sqrt(B.fused([dict(z=Nv, u=tn, w=Nv, a=1.0, b=-beta)], [(Nv, Nv)])[0])
                else:
                    B.fused([dict(z=Nv, u=tn, w=Nv, a=1.0, b=-beta)])
                    B.apply_callable(N, Nv, v)
                    alpha = sqrt(B.fused([], [(v, Nv)])[0])
                if alpha > 0:
                    ops = [dict(z=v, u=v, a=alpha, a_div=True)]
                    if N is not None:
                        ops.append(dict(z=Nv, u=Nv, a=alpha, a_div=True))
                    B.fused(ops)
            beta_hat = c * beta                                        # craigmr.py:159-176
            gamma = s * beta
            delta = sqrt(gamma ** 2 + 1)
            alpha_hat = sqrt(alpha ** 2 + delta ** 2)
            c = alpha / alpha_hat
            s = delta / alpha_hat
            rho = sqrt(alpha_tilde ** 2 + beta_hat ** 2)
            c_hat = alpha_tilde / rho
            s_hat = beta_hat / rho
            theta_old = theta
            theta = s_hat * alpha_hat
            alpha_tilde = -c_hat * alpha_hat
            zeta = c_hat * zeta_hat
            zeta_hat = s_hat * zeta_hat
            xNrgNorm2 += zeta * zeta
            print(itn, xNrgNorm2)                                       # (sic) craigmr.py:190
            # dbar = (d - theta dbar)/rho ; d = (u - beta_hat d)/alpha_hat ; x += zeta dbar   (:178,191-192)
            B.fused([dict(z=dbar, u=d, w=dbar, a=1.0, b=-theta_old), dict(z=dbar, u=dbar, a=rho, a_div=True),
                     dict(z=d, u=u, w=d, a=1.0, b=-beta_hat), dict(z=d, u=d, a=alpha_hat, a_div=True)])
            B.fused([dict(z=x, u=x, w=dbar, a=1.0, b=zeta)])
            if store_iterates:
                self.iterates.append(x.download())
            if store_resids:
                self.norms.append(xNrgNorm2)
                self.normal_eqns_resids.append(abs(zeta))
            dErr[itn % window] = zeta
            if itn > window:
                trncDirErr = np.linalg.norm(dErr)
                xNrgNorm = sqrt(xNrgNorm2)
                self.dir_errors_window.append(trncDirErr / xNrgNorm)
                if trncDirErr < etol * xNrgNorm:
                    istop = 8
            if itn >= itnlim:
                istop = 7
            if istop > 0:
                break
        if show:
            print(" ")
            print("CRAIG-MR finished")
            print(self.msg[istop])
            print(" ")
            print("xNrgNorm2 = %7.1e   trnDirErr = %7.1e" % (xNrgNorm2, trncDirErr))
            print(" ")
        if istop == 0:
            self.status = "solution is zero"
        if istop in [1, 2, 4, 5]:
            self.status = "residual small"
        if istop in [3, 6]:
            self.status = "ill-conditioned operator"
        if istop == 7:
            self.status = "max iterations"
        if istop == 8:
            self.status = "direct error small"
        self.optimal = istop in [1, 2, 4, 5, 8]
        self.x = self.bestSolution = x.download()
        self.istop = istop
        self.itn = itn
        self.nMatvec = 2 * itn
        return
