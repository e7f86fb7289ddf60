"""LSMR on device vectors (reference: pykrylov/lls/lsmr.py:28-518, Fong & Saunders).

Same call surface as the reference's ``LSMRFramework.solve`` -- including its quirk of
*returning* ``(x, istop, itn, normr, normar, normA, condA, normx)`` and setting only
``self.x`` (lsmr.py:491-492).  The Golub-Kahan step and the ``h / hbar / x`` recurrences
(lsmr.py:309-351) run as CUDA kernels (CSR SpMV with A and the device-built A^T, fused
multi-AXPY+dot launches); the rotations and norm estimates stay on the host.
"""
from math import sqrt

import numpy as np

from ..generic import KrylovMethod
from .. import _engine

__docformat__ = "restructuredtext"


def sign(a):
    return -1 if a < 0 else 1


def symOrtho(a, b):
    """Stable Givens rotation (S.-C. Choi's thesis; reference lsmr.py:500-518)."""
    if b == 0:
        return sign(a), 0, abs(a)
    elif a == 0:
        return 0, sign(b), abs(b)
    elif abs(b) > abs(a):
        tau = a / b
        s = sign(b) / sqrt(1 + tau * tau)
        c = s * tau
        r = b / s
    else:
        tau = b / a
        c = sign(a) / sqrt(1 + tau * tau)
        s = c * tau
        r = a / c
    return c, s, r


class LSMRFramework(KrylovMethod):
    """LSMR: ``A x = b`` or ``min |b - A x|_2`` (optionally damped) for any m x n ``A``."""

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
        self.name = "Least-Squares Minimum Residual"
        self.acronym = "LSMR"
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
        self.norms = []
        self.dir_errors_window = []
        self.iterates = []

    def solve(self, b, damp=0.0, atol=1e-9, btol=1e-9, conlim=1e8, M=None, N=None, itnlim=None,
              show=False, **kwargs):
        etol = kwargs.get("etol", 1.0e-6)
        store_resids = kwargs.get("store_resids", False)
        store_iterates = kwargs.get("store_iterates", False)
        window = kwargs.get("window", 5)
        self.resids, self.normal_eqns_resids, self.norms = [], [], []
        self.dir_errors_window, self.iterates = [], []
        A = self.A
        b = np.asarray(b, dtype=np.float64).squeeze()
        msg = self.msg
        hdg1 = "   itn      x(1)       norm r    norm A'r"
        hdg2 = " compatible   LS      norm A   cond A"
        pfreq, pcount = 20, 0
        m, n = A.shape
        minDim = min([m, n])
        if itnlim is None:
            itnlim = minDim
        if show:
            print(" ")
            print("LSMR            Least-squares solution of  Ax = b")
            print("The matrix A has %8g rows  and %8g cols" % (m, n))
            print("damp = %20.14e" % (damp))
            print("atol = %8.2e                 conlim = %8.2e" % (atol, conlim))
            print("btol = %8.2e               itnlim = %8g" % (btol, itnlim))

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
        zetabar = alpha * beta
        alphabar = alpha
        rho = rhobar = cbar = 1
        sbar = 0
        h, hbar, x = B.vec_n(n), B.vec_n(n), B.vec_n(n)
        B.fused([dict(z=h, u=v, a=1.0)])
        if store_iterates:
            self.iterates.append(x.download())
        betadd, betad, rhodold, tautildeold, thetatilde, zeta, d = beta, 0, 1, 0, 0, 0, 0
        normA2 = alpha * alpha
        maxrbar, minrbar = 0, 1e+100
        normA = sqrt(normA2)
        condA = 1
        normx = 0
        xNrgNorm2 = 0
        dErr = np.zeros(window)
        trncDirErr = 0
        normb = beta
        istop = 0
        ctol = 0
        if conlim > 0:
            ctol = 1 / conlim
        normr = beta
        normar = alpha * beta
        if normar == 0:
            if show:
                print(msg[0])
            return x.download(), istop, itn, normr, normar, normA, condA, normx   # (self.x not set: lsmr.py:287)
        if show:
            print(" ")
            print(hdg1, hdg2)
            print("".join(["%6g %12.5e" % (itn, 0.0), " %10.3e %10.3e" % (normr, normar),
                           "  %8.1e %8.1e" % (1, alpha / beta)]))
        if store_resids:
            self.resids.append(normr)
            self.normal_eqns_resids.append(normar)

        csr = _engine.plane_csr(A)
        on_device = (csr is not None and M is None and N is None and not show and not store_iterates and itnlim > 0)
        if on_device:
            # device-resident loop (scalar plane in csrc/lls.cu, see lsqr.py)
            from ..device import ScalarPlane as SL
            loop = _engine.PlaneLoop(B.ctx, "lsmr")
            loop.P.setup(dict(alpha=alpha, beta=beta, zetabar=zetabar, alphabar=alphabar, rho=rho, rhobar=rhobar,
                              cbar=cbar, sbar=sbar, betadd=betadd, betad=betad, rhodold=rhodold,
                              tautildeold=tautildeold, thetatilde=thetatilde, zeta=zeta, d=d, normA2=normA2,
                              maxrbar=maxrbar, minrbar=minrbar, normA=normA, condA=condA, normx=normx,
                              normb=normb, normr=normr, normar=normar),
                         window=window, itnlim=itnlim, damp=damp, atol=atol, btol=btol, ctol=ctol, etol=etol)

            def trip():
                # Mu = A v - alpha Mu, |Mu|^2 and phase 1 (beta, |A|) in one launch; likewise A'u below
                loop.spmv_ops(csr, v, dict(z=Mu, w=Mu, a=1.0, b_slot=SL.ALPHA, b_neg=1), step=1)
                loop.ops([dict(z=u, u=u, a_slot=SL.U_DIV, a_div=True)])
                loop.spmv_ops(csr, u, dict(z=Nv, w=Nv, a_slot=SL.NV_A, b_slot=SL.NV_B), step=2, trans=True)
                loop.ops([dict(z=v, u=v, a_slot=SL.V_DIV, a_div=True), dict(z=hbar, u=h, w=hbar, a=1.0, b_slot=SL.C0),
                          dict(z=x, u=x, w=hbar, a=1.0, b_slot=SL.C1), dict(z=h, u=v, w=h, a=1.0, b_slot=SL.C2)],
                         [(x, x)], step=3)

            def replay(st_, sc_, hist):
                for nr, nar, nrg, direrr in hist:
                    if store_resids:
                        self.norms.append(nrg)
                        self.resids.append(nr)
                        self.normal_eqns_resids.append(nar)
                    if direrr == direrr:
                        self.dir_errors_window.append(direrr)

            if not csr.symmetric:
                csr.build_transpose()
            st_, sc = loop.run(trip, self.check_interval, replay)
            itn, istop = int(st_.itn), int(st_.istop)
            normr, normar, normA, condA, normx = sc["normr"], sc["normar"], sc["normA"], sc["condA"], sc["normx"]
            xNrgNorm2 = sc["xNrgNorm2"]
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
            chat, shat, alphahat = symOrtho(alphabar, damp)                   # lsmr.py:337-353
            rhoold = rho
            c, s, rho = symOrtho(alphahat, beta)
            thetanew = s * alpha
            alphabar = c * alpha
            rhobarold = rhobar
            zetaold = zeta
            thetabar = sbar * rho
            rhotemp = cbar * rho
            cbar, sbar, rhobar = symOrtho(cbar * rho, thetanew)
            zeta = cbar * zetabar
            zetabar = -sbar * zetabar
            # v /= alpha ; hbar = h - c1 hbar ; x += c2 hbar ; h = v - c3 h ; |x|^2  -- one launch
            ops = []
            if scale_v:
                ops.append(dict(z=v, u=v, a=alpha, a_div=True))
                if N is not None:
                    B.fused([dict(z=Nv, u=Nv, a=alpha, a_div=True)])
            ops += [dict(z=hbar, u=h, w=hbar, a=1.0, b=-(thetabar * rho / (rhoold * rhobarold))),
                    dict(z=x, u=x, w=hbar, a=1.0, b=(zeta / (rho * rhobar))),
                    dict(z=h, u=v, w=h, a=1.0, b=-(thetanew / rho))]
            xx = B.fused(ops, [(x, x)])[0]
            if store_iterates:
                self.iterates.append(x.download())
            xNrgNorm2 += zeta * zeta
            dErr[itn % window] = zeta
            if itn > window:
                trncDirErr = np.linalg.norm(dErr)
                xNrgNorm = sqrt(xNrgNorm2)
                self.dir_errors_window.append(trncDirErr / xNrgNorm)
                if trncDirErr < etol * xNrgNorm:
                    istop = 8
            betaacute = chat * betadd                                        # estimate of ||r||, :368-392
            betacheck = -shat * betadd
            betahat = c * betaacute
            betadd = -s * betaacute
            thetatildeold = thetatilde
            ctildeold, stildeold, rhotildeold = symOrtho(rhodold, thetabar)
            thetatilde = stildeold * rhobar
            rhodold = ctildeold * rhobar
            betad = -stildeold * betad + ctildeold * betahat
            tautildeold = (zetaold - thetatildeold * tautildeold) / rhotildeold
            taud = (zeta - thetatilde * tautildeold) / rhodold
            d = d + betacheck * betacheck
            normr = sqrt(d + (betad - taud) ** 2 + betadd * betadd)
            normA2 = normA2 + beta * beta                                    # ||A||, cond(A), :395-403
            normA = sqrt(normA2)
            normA2 = normA2 + alpha * alpha
            maxrbar = max(maxrbar, rhobarold)
            if itn > 1:
                minrbar = min(minrbar, rhobarold)
            condA = max(maxrbar, rhotemp) / min(minrbar, rhotemp)
            normar = abs(zetabar)
            normx = np.sqrt(xx)
            test1 = normr / normb
            test2 = normar / (normA * normr)
            test3 = 1 / condA
            t1 = test1 / (1 + normA * normx / normb)
            rtol = btol + atol * normA * normx / normb
            if store_resids:
                self.norms.append(xNrgNorm2)
                self.resids.append(normr)
                self.normal_eqns_resids.append(normar)
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
            if show and (n <= 40 or itn <= 10 or itn >= itnlim - 10 or itn % 10 == 0 or test3 <= 1.1 * ctol
                         or test2 <= 1.1 * atol or test1 <= 1.1 * rtol or istop != 0):
                if pcount >= pfreq:
                    pcount = 0
                    print(" ")
                    print(hdg1, hdg2)
                pcount += 1
                print("".join(["%6g %12.5e" % (itn, x.peek(0)), " %10.3e %10.3e" % (normr, normar),
                               "  %8.1e %8.1e" % (test1, test2), " %8.1e %8.1e" % (normA, condA)]))
            if istop > 0:
                break
        if show:
            print(" ")
            print("LSMR finished")
            print(msg[istop])
            print("istop =%8g    normr =%8.1e" % (istop, normr), "    normA =%8.1e    normAr =%8.1e" % (normA, normar))
            print("itn   =%8g    condA =%8.1e" % (itn, condA), "    normx =%8.1e" % (normx))
            print("Estimated energy norm of x: %7.1e" % sqrt(xNrgNorm2))
        self.x = x.download()
        return self.x, istop, itn, normr, normar, normA, condA, normx
