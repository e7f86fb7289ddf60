"""CRAIG on device vectors (reference: pykrylov/lls/craig.py:30-525, Arioli & Orban's
generalised CRAIG for symmetric quasi-definite systems).

Same keyword contract and attribute set as the reference's ``CRAIGFramework``; like the
reference, most of LSQR's norm estimates are absent (they are commented out there).  The
Golub-Kahan step and the primal / dual updates (craig.py:300-362) run as CUDA kernels
(CSR SpMV with A and the device-built A^T, fused multi-AXPY+dot launches); the rotations
stay on the host.
"""
from math import sqrt

import numpy as np

from ..generic import KrylovMethod
from .. import _engine
from .lsqr import normof2

__docformat__ = "restructuredtext"


class CRAIGFramework(KrylovMethod):
    r"""Generalised CRAIG: ``min |b - A x|^2_D + |x|^2_N``, i.e. the SQD system
    ``[M A; A' -N] [r; x] = [b; 0]`` with ``M = inv(D)``."""

    msg = ["The exact solution is  x = 0                              ",
           "Ax - b is small enough, given atol, btol                  ",
           "The least-squares solution is good enough, given atol     ",
           "The estimate of cond(Abar) has exceeded conlim            ",
           "Ax - b is small enough for this machine                   ",
           "The least-squares solution is good enough for this machine",
           "Cond(Abar) seems to be too large for this machine         ",
           "The iteration limit has been reached                      ",
           "The truncated direct error is small enough, given etol    "]

    def __init__(self, A, **kwargs):
        KrylovMethod.__init__(self, A, **kwargs)
        self.name = "CRAIG's Method for Least Squares"
        self.acronym = "CRAIG"
        self.prefix = self.acronym + ": "
        self.A = A
        self.x = None
        self.var = None
        self.itn = 0
        self.istop = 0
        self.Anorm = self.Acond = self.Arnorm = self.xnorm = 0.
        self.r1norm = self.r2norm = 0.
        self.optimal = False
        self.norms, self.resids, self.normal_eqns_resids = [], [], []
        self.dir_errors_p_window, self.dir_errors_d_window = [], []
        self.iterates_p, self.iterates_d = [], []

    def solve(self, rhs, itnlim=0, damp=0.0, M=None, N=None, atol=1.0e-9, btol=1.0e-9,
              conlim=1.0e+8, show=False, wantvar=False, **kwargs):
        etol = kwargs.get("etol", 1.0e-6)
        store_resids = kwargs.get("store_resids", False)
        store_iterates = kwargs.get("store_iterates", False)
        window = kwargs.get("window", 5)
        self.norms, self.resids, self.normal_eqns_resids = [], [], []
        self.dir_errors_p_window, self.dir_errors_d_window = [], []
        self.iterates_p, self.iterates_d = [], []
        A = self.A
        m, n = A.shape
        if itnlim == 0:
            itnlim = 3 * n
        if wantvar:
            raise NotImplementedError("wantvar=True fails in the reference itself (zeros(n,1), craig.py:178)")
        itn = istop = 0
        if conlim > 0.0:
            self.ctol = 1.0 / conlim
        if show:
            print(" ")
            print("CRAIG           Least-squares solution of  Ax = b")
            print("The matrix A has %8d rows and %8d cols" % (m, n))
            print("damp = %20.14e     wantvar = %-5s" % (damp, repr(wantvar)))
            print("atol = %8.2e                 conlim = %8.2e" % (atol, conlim))
            print("btol = %8.2e                 itnlim = %8g" % (btol, itnlim))

        B = _engine.HostBridge(n, self.context, A)
        r, x = B.vec_n(m), B.vec_n(n)
        rNrgNorm2 = xNrgNorm2 = 0.0
        dErr = np.zeros(window)
        trncDirErr = 0
        Mu = B.vec_n(m, np.asarray(rhs[:m], dtype=np.float64))
        u = Mu if M is None else B.apply_callable(M, Mu, B.vec_n(m))
        Nv = B.vec_n(n)
        v = Nv if N is None else B.vec_n(n)
        tm, tn = B.vec_n(m), B.vec_n(n)
        d, w, wbar = B.vec_n(m), B.vec_n(n), B.vec_n(n)
        alpha = 0.
        beta = sqrt(B.fused([], [(u, Mu)])[0])
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
        x_is_zero = False
        if beta == 0.0:
            if show:
                print(self.msg[0])
            x_is_zero = True
            istop = 0
        bnorm = beta
        delta = 1.
        rho = normof2(alpha, 1)
        tau = beta / rho                                               # dual variables, craig.py:249-252
        B.fused([dict(z=d, u=u, a=rho, a_div=True), dict(z=r, u=d, a=tau)])
        rnorm = tau * tau
        c = alpha / rho                                                # primal variables, :255-263
        s = 1 / rho
        zeta = s * beta
        eta = c * zeta
        xi = s * zeta
        B.fused([dict(z=w, u=v, a=c), dict(z=wbar, u=v, a=s), dict(z=x, u=w, a=zeta)])
        xnorm = eta * eta
        r1norm = xi * xi
        r2norm = rnorm
        Arnorm = 0.0
        if store_iterates:
            self.iterates_p.append(x.download())
            self.iterates_d.append(r.download())
        if store_resids:
            self.norms.append(xNrgNorm2)
            self.resids.append(r2norm)

        csr = _engine.plane_csr(A)
        on_device = (csr is not None and M is None and N is None and not show and not store_iterates
                     and not x_is_zero and itnlim > 0)
        if on_device:
            # device-resident loop (scalar plane in csrc/lls.cu, see lsqr.py)
            from ..device import ScalarPlane as SL
            loop = _engine.PlaneLoop(B.ctx, "craig")
            loop.P.setup(dict(alpha=alpha, beta=beta, c=c, s=s, tau=tau, zeta=zeta, eta=eta, xi=xi, rnorm=rnorm,
                              xnorm=xnorm, r1norm=r1norm, r2norm=r2norm, Arnorm=Arnorm, rNrgNorm2=rNrgNorm2,
                              xNrgNorm2=xNrgNorm2, bnorm=bnorm),
                         window=window, itnlim=itnlim, btol=btol, etol=etol)

            def trip():
                # Mu = A v - alpha Mu, |Mu|^2 and phase 1 (beta, |A|) in one launch; likewise A'u below
                loop.spmv_ops(csr, v, dict(z=Mu, w=Mu, a=1.0, b_slot=SL.ALPHA, b_neg=1), step=1)
                loop.ops([dict(z=u, u=u, a_slot=SL.U_DIV, a_div=True)])
                loop.spmv_ops(csr, u, dict(z=Nv, w=Nv, a_slot=SL.NV_A, b_slot=SL.NV_B), step=2, trans=True)
                loop.ops([dict(z=v, u=v, a_slot=SL.V_DIV, a_div=True)])
                loop.ops([dict(z=d, u=u, w=d, a=1.0, b_slot=SL.C0, b_neg=1), dict(z=d, u=d, a_slot=SL.C1, a_div=True),
                          dict(z=r, u=r, w=d, a=1.0, b_slot=SL.C2)])
                loop.ops([dict(z=wbar, u=wbar, a_slot=SL.C3), dict(z=w, u=v, w=wbar, a_slot=SL.C4, b_slot=SL.C5),
                          dict(z=wbar, u=wbar, w=v, a_slot=SL.C4, a_neg=1, b_slot=SL.C5),
                          dict(z=x, u=x, w=w, a=1.0, b_slot=SL.C6)])
                loop.P.step(9)                              # latch `done` behind this trip's updates

            def replay(st_, sc_, hist):
                for r2, arn, nrg, direrr in hist:
                    if store_resids:
                        self.norms.append(nrg)
                        self.resids.append(r2)
                        self.normal_eqns_resids.append(arn)
                    if direrr == direrr:
                        self.dir_errors_d_window.append(direrr)

            if not csr.symmetric:
                csr.build_transpose()
            st_, sc = loop.run(trip, self.check_interval, replay)
            itn, istop = int(st_.itn), int(st_.istop)
            r1norm, r2norm, Arnorm, xnorm = sc["r1norm"], sc["r2norm"], sc["Arnorm"], sc["xnorm"]
            xNrgNorm2, trncDirErr = sc["xNrgNorm2"], sc["trncDirErr"]
            A._nMatvec += 2 * itn

        while itn < itnlim and not x_is_zero and not on_device:
            itn += 1
            B.apply(A, v, tm)
            if M is None:
                beta = sqrt(B.fused([dict(z=Mu, u=tm, w=Mu, a=1.0, b=-alpha)], [(Mu, Mu)])[0])
            else:
                B.fused([dict(z=Mu, u=tm, w=Mu, a=1.0, b=-alpha)])
                B.apply_callable(M, Mu, u)
                beta = sqrt(B.fused([], [(u, Mu)])[0])
            Arnorm = abs(alpha * beta * s * zeta)                      # :314
            scale_v = False
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
                scale_v = alpha > 0
            beta_hat = c * beta                                        # rotations, :336-347
            gamma = s * beta
            delta = normof2(gamma, 1)
            s2 = gamma / delta
            alpha_hat = normof2(alpha, delta)
            c = alpha / alpha_hat
            s = delta / alpha_hat
            tau = -beta_hat * tau / alpha_hat
            zeta = -beta_hat * zeta / alpha_hat
            eta = c * zeta
            xi = s * zeta
            # v /= alpha ; d = (u - beta_hat d)/alpha_hat ; r += tau d                       (:350-352)
            if scale_v:                       # (length-n vectors: their own launch)
                ops = [dict(z=v, u=v, a=alpha, a_div=True)]
                if N is not None:
                    ops.append(dict(z=Nv, u=Nv, a=alpha, a_div=True))
                B.fused(ops)
            B.fused([dict(z=d, u=u, w=d, a=1.0, b=-beta_hat), dict(z=d, u=d, a=alpha_hat, a_div=True),
                     dict(z=r, u=r, w=d, a=1.0, b=tau)])
            # wbar *= s2 ; w = c v + s wbar ; wbar = -c wbar + s v ; x += zeta w            (:360-365)
            B.fused([dict(z=wbar, u=wbar, a=s2), dict(z=w, u=v, w=wbar, a=c, b=s),
                     dict(z=wbar, u=wbar, w=v, a=-c, b=s), dict(z=x, u=x, w=w, a=1.0, b=zeta)])
            if store_iterates:
                self.iterates_p.append(x.download())
                self.iterates_d.append(r.download())
            rNrgNorm2 += tau * tau
            xNrgNorm2 += zeta * zeta
            dErr[itn % window] = tau
            if itn > window:
                trncDirErr = np.linalg.norm(dErr)
                rNrgNorm = sqrt(rNrgNorm2)
                self.dir_errors_d_window.append(trncDirErr / rNrgNorm)
                if trncDirErr < etol * rNrgNorm:
                    istop = 8
            rnorm += tau * tau
            xnorm += eta * eta
            r1norm += xi * xi
            r2norm = rnorm
            test1 = sqrt(rnorm) / bnorm
            t1 = test1
            rtol = btol
            if store_resids:
                self.norms.append(xNrgNorm2)
                self.resids.append(r2norm)
                self.normal_eqns_resids.append(Arnorm)
            if itn >= itnlim:
                istop = 7
            if 1 + t1 <= 1:
                istop = 4
            if test1 <= rtol:
                istop = 1
            if istop > 0:
                break

        if show:
            print(" ")
            print("CRAIG finished")
            print(self.msg[istop])
            print(" ")
            print("istop =%8g   r1norm =%8.1e" % (istop, sqrt(r1norm)))
            print("itn   =%8g   r2norm =%8.1e" % (itn, sqrt(r2norm)))
            print("                  bnorm  =%8.1e" % bnorm)
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
        self.r = r.download()
        self.istop = istop
        self.itn = itn
        self.nMatvec = 2 * itn
        self.r1norm = sqrt(r1norm)
        self.r2norm = sqrt(r2norm)
        self.Arnorm = Arnorm
        self.xnorm = xnorm
        return
