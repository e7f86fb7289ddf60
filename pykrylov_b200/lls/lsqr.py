"""LSQR on device vectors (reference: pykrylov/lls/lsqr.py:26-455).

Same class surface and keyword contract as the reference's ``LSQRFramework``.  The
Golub-Kahan step (lsqr.py:243-275) runs as CUDA kernels -- ``A v`` / ``A^T u`` are
CSR SpMVs on the operator and its device-built transpose, the vector recurrences
and norms are fused multi-AXPY+dot launches -- while the scalar plane rotations
and stopping tests (lsqr.py:277-390, a few dozen flops) stay on the host, which
reads three inner products per iteration.  Opaque operators / preconditioners
``M``, ``N`` are applied through the host bridge.
"""
from math import sqrt

import numpy as np

from ..generic import KrylovMethod
from .. import _engine

__docformat__ = "restructuredtext"

inf = float("inf")


def normof2(x, y):
    return sqrt(x * x + y * y)


def normof4(x1, x2, x3, x4):
    return sqrt(x1 * x1 + x2 * x2 + x3 * x3 + x4 * x4)


class LSQRFramework(KrylovMethod):
    r"""LSQR: ``A x = b``, ``min |b - A x|`` or, with ``damp > 0``,
    ``min |b - A x|^2 + damp^2 |x|^2`` for an (m x n) operator ``A`` providing
    ``A * x`` and ``A.T * u`` (Paige & Saunders 1982)."""

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
        self.name = "Least-Squares QR"
        self.acronym = "LSQR"
        self.prefix = self.acronym + ": "
        self.A = A
        self.x = None
        self.var = None
        self.itn = 0
        self.istop = 0
        self.Anorm = self.Acond = self.Arnorm = self.xnorm = 0.
        self.r1norm = self.r2norm = 0.
        self.optimal = False
        self.resids = []
        self.normal_eqns_resids = []
        self.dir_errors_window = []
        self.error_upper_bound = []
        self.iterates = []

    def solve(self, rhs, itnlim=0, damp=0.0, M=None, N=None, atol=1.0e-9, btol=1.0e-9,
              conlim=1.0e+8, show=False, wantvar=False, **kwargs):
        """Keywords as in lsqr.py:85-148 (``etol`` 1e-6, ``store_resids``,
        ``store_iterates``, ``window`` 5).  Results are left in the attributes
        ``x, istop, itn, r1norm, r2norm, Anorm, Acond, Arnorm, xnorm, var``."""
        etol = kwargs.get("etol", 1.0e-6)
        store_resids = kwargs.get("store_resids", False)
        store_iterates = kwargs.get("store_iterates", False)
        window = kwargs.get("window", 5)
        self.resids, self.normal_eqns_resids = [], []
        self.dir_errors_window, self.iterates = [], []

        A = self.A
        m, n = A.shape
        if itnlim == 0:
            itnlim = 3 * n
        if wantvar:
            raise NotImplementedError("wantvar=True fails in the reference itself (zeros(n,1), lsqr.py:155)")
        var = None
        dampsq = damp * damp
        itn = istop = 0
        ctol = 0.0
        if conlim > 0.0:
            self.ctol = 1.0 / conlim          # (sic) the local ctol stays 0.0, lsqr.py:162-163
        Anorm = Acond = 0.
        z = xnorm = xxnorm = ddnorm = res2 = 0.
        cs2, sn2 = -1., 0.

        if show:
            print(" ")
            print("LSQR            Least-squares solution of  Ax = b")
            print("The matrix A has %8d rows and %8d cols" % (m, n))
            print("damp = %20.14e     wantvar = %-5s" % (damp, repr(wantvar)))
            print("atol = %8.2e                 conlim = %8.2e" % (atol, conlim))
            print("btol = %8.2e                 itnlim = %8g" % (btol, itnlim))

        B = _engine.HostBridge(n, self.context, A)
        x = B.vec_n(n)
        xNrgNorm2 = 0.0
        dErr = np.zeros(window)
        trncDirErr = 0
        if store_iterates:
            self.iterates.append(x.download())

        # first vectors of the bidiagonalisation: beta M u = b, alpha N v = A'u
        Mu = B.vec_n(m, np.asarray(rhs[:m], dtype=np.float64))
        u = Mu if M is None else B.apply_callable(M, Mu, B.vec_n(m))
        tm, tn = B.vec_n(m), B.vec_n(n)           # SpMV outputs
        Nv = B.vec_n(n)
        v = Nv if N is None else B.vec_n(n)
        w, dk = B.vec_n(n), B.vec_n(n)
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
            ops.append(dict(z=w, u=v, a=1.0))
            B.fused(ops)

        x_is_zero = False
        Arnorm = alpha * beta
        if Arnorm == 0.0:
            if show:
                print(self.msg[0])
            x_is_zero = True
            istop = 0
        rhobar, phibar, bnorm = alpha, beta, beta
        rnorm = r1norm = r2norm = beta
        if show:
            print(" ")
            print("   Itn      x(1)       r1norm     r2norm  Compatible   LS      Norm A   Cond A")
            test2 = alpha / beta if not x_is_zero else 1.0
            print("%6g %12.5e %10.3e %10.3e  %8.1e %8.1e" % (itn, 0.0, r1norm, r2norm, 1.0, test2))
        if store_resids:
            self.resids.append(r2norm)
            self.normal_eqns_resids.append(Arnorm)

        csr = _engine.plane_csr(A)
        on_device = (csr is not None and M is None and N is None and not show and not store_iterates
                     and not x_is_zero and itnlim > 0)
        if on_device:
            # device-resident loop: the scalar plane (rotations, norms, stopping tests) runs in
            # csrc/lls.cu, the host enqueues whole trips and reads one status block per check interval
            from ..device import ScalarPlane as SL
            loop = _engine.PlaneLoop(B.ctx, "lsqr")
            loop.P.setup(dict(alpha=alpha, beta=beta, rhobar=rhobar, phibar=phibar, bnorm=bnorm, cs2=cs2, sn2=sn2,
                              rnorm=rnorm, r1norm=r1norm, r2norm=r2norm, Arnorm=Arnorm),
                         window=window, itnlim=itnlim, damp=damp, atol=atol, btol=btol, ctol=ctol, etol=etol)

            def trip():
                # Mu = A v - alpha Mu, |Mu|^2 and phase 1 (beta, |A|) in one launch; likewise A'u below
                loop.spmv_ops(csr, v, dict(z=Mu, w=Mu, a=1.0, b_slot=SL.ALPHA, b_neg=1), step=1)
                loop.ops([dict(z=u, u=u, a_slot=SL.U_DIV, a_div=True)])
                loop.spmv_ops(csr, u, dict(z=Nv, w=Nv, a_slot=SL.NV_A, b_slot=SL.NV_B), step=2, trans=True)
                loop.ops([dict(z=v, u=v, a_slot=SL.V_DIV, a_div=True), dict(z=dk, u=w, a_slot=SL.C0),
                          dict(z=x, u=x, w=w, a=1.0, b_slot=SL.C1), dict(z=w, u=w, w=v, a_slot=SL.C2, b=1.0)],
                         [(dk, dk)], step=3)                                             # norms, stopping tests

            def replay(st_, sc_, hist):
                for r2, arn, direrr, _ in hist:
                    if store_resids:
                        self.resids.append(r2)
                        self.normal_eqns_resids.append(arn)
                    if direrr == direrr:
                        self.dir_errors_window.append(direrr)

            if not csr.symmetric:
                csr.build_transpose()
            st_, sc = loop.run(trip, self.check_interval, replay)
            itn, istop = int(st_.itn), int(st_.istop)
            Anorm, Acond, Arnorm, xnorm = sc["Anorm"], sc["Acond"], sc["Arnorm"], sc["xnorm"]
            r1norm, r2norm = sc["r1norm"], sc["r2norm"]
            xNrgNorm2, trncDirErr = sc["xNrgNorm2"], sc["trncDirErr"]
            A._nMatvec += 2 * itn

        while itn < itnlim and not x_is_zero and not on_device:
            itn += 1
            # beta M u = A v - alpha M u
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
                Anorm = normof4(Anorm, alpha, beta, damp)
                # alpha N v = A'u - beta N v
                B.apply(A, u, tn, trans=True)
                if N is None:
                    alpha = sqrt(B.fused([dict(z=Nv, u=tn, w=Nv, a=1.0, b=-beta)], [(Nv, Nv)])[0])
                else:
                    B.fused([dict(z=Nv, u=tn, w=Nv, a=1.0, b=-beta)])
                    B.apply_callable(N, Nv, v)
                    alpha = sqrt(B.fused([], [(v, Nv)])[0])
            # plane rotations (lsqr.py:277-296)
            rhobar1 = normof2(rhobar, damp)
            cs1 = rhobar / rhobar1
            sn1 = damp / rhobar1
            psi = sn1 * phibar
            phibar = cs1 * phibar
            rho = normof2(rhobar1, beta)
            cs = rhobar1 / rho
            sn = beta / rho
            theta = sn * alpha
            rhobar = -cs * alpha
            phi = cs * phibar
            phibar = sn * phibar
            tau = sn * phi
            # v /= alpha ; dk = w/rho ; x += t1 w ; w = t2 w + v ; |dk|^2   -- one launch
            t1 = phi / rho
            t2 = -theta / rho
            ops = []
            if beta > 0 and alpha > 0:
                ops.append(dict(z=v, u=v, a=alpha, a_div=True))
                if N is not None:
                    B.fused([dict(z=Nv, u=Nv, a=alpha, a_div=True)])
            ops += [dict(z=dk, u=w, a=1.0 / rho), dict(z=x, u=x, w=w, a=1.0, b=t1),
                    dict(z=w, u=w, w=v, a=t2, b=1.0)]
            dk2 = B.fused(ops, [(dk, dk)])[0]
            ddnorm = ddnorm + np.sqrt(dk2) ** 2
            if store_iterates:
                self.iterates.append(x.download())
            xNrgNorm2 += phi * phi
            dErr[itn % window] = phi
            if itn > window:
                trncDirErr = np.linalg.norm(dErr)
                xNrgNorm = sqrt(xNrgNorm2)
                self.dir_errors_window.append(trncDirErr / xNrgNorm)
                if trncDirErr < etol * xNrgNorm:
                    istop = 8
            # estimate norm(x) (lsqr.py:324-332)
            delta = sn2 * rho
            gambar = -cs2 * rho
            rhs_ = phi - delta * z
            zbar = rhs_ / gambar
            xnorm = sqrt(xxnorm + zbar ** 2)
            gamma = normof2(gambar, theta)
            cs2 = gambar / gamma
            sn2 = theta / gamma
            z = rhs_ / gamma
            xxnorm += z * z
            # convergence tests (lsqr.py:338-390)
            Acond = Anorm * sqrt(ddnorm)
            res1 = phibar ** 2
            res2 = res2 + psi ** 2
            rnorm = sqrt(res1 + res2)
            Arnorm = alpha * abs(tau)
            r1sq = rnorm ** 2 - dampsq * xxnorm
            r1norm = sqrt(abs(r1sq))
            if r1sq < 0:
                r1norm = -r1norm
            r2norm = rnorm
            test1 = rnorm / bnorm
            test2 = inf if (Anorm == 0. or rnorm == 0.) else Arnorm / (Anorm * rnorm)
            test3 = inf if Acond == 0.0 else 1.0 / Acond
            t1 = test1 / (1 + Anorm * xnorm / bnorm)
            rtol = btol + atol * Anorm * xnorm / bnorm
            if store_resids:
                self.resids.append(r2norm)
                self.normal_eqns_resids.append(Arnorm)
            if itn >= itnlim:
                istop = 7
            if 1 + test3 <= 1:
                istop = 6
            if 1 + test2 <= 1:
                istop = 5
            if 1 + t1 <= 1:
                istop = 4
            if test3 <= ctol:
                istop = 3
            if test2 <= atol:
                istop = 2
            if test1 <= rtol:
                istop = 1
            if show and (n <= 40 or itn <= 10 or itn >= itnlim - 10 or itn % 10 == 0 or test3 <= 2 * ctol
                         or test2 <= 10 * atol or test1 <= 10 * rtol or istop != 0):
                print("%6g %12.5e %10.3e %10.3e  %8.1e %8.1e %8.1e %8.1e"
                      % (itn, x.download()[0], r1norm, r2norm, test1, test2, Anorm, Acond))
            if istop > 0:
                break

        if show:
            print(" ")
            print("LSQR finished")
            print(self.msg[istop])
            print(" ")
            print("istop =%8g   r1norm =%8.1e   Anorm =%8.1e   Arnorm =%8.1e" % (istop, r1norm, Anorm, Arnorm))
            print("itn   =%8g   r2norm =%8.1e   Acond =%8.1e   xnorm  =%8.1e" % (itn, r2norm, Acond, xnorm))
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
        self.istop = istop
        self.itn = itn
        self.nMatvec = 2 * itn
        self.r1norm = r1norm
        self.r2norm = r2norm
        self.residNorm = r2norm
        self.Anorm = Anorm
        self.Acond = Acond
        self.Arnorm = Arnorm
        self.xnorm = xnorm
        self.var = var
        return


class LSQR(LSQRFramework):
    """Convenience alias (the reference exports only ``LSQRFramework``)."""
