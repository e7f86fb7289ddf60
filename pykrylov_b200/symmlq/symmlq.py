"""SYMMLQ on device vectors (reference: pykrylov/symmlq/symmlq.py:16-420).

Same keyword contract as the reference's ``Symmlq.solve`` (``matvec_max`` 2n+2,
``rtol`` 1e-9, ``check``, ``shift``, ``store_iterates``).  The Lanczos step
(symmlq.py:300-313) and the x / w updates (:335-336) run as CUDA kernels (CSR SpMV
+ fused multi-AXPY+dot launches); the plane rotations and tests stay on the host,
which reads two inner products per iteration.

The reference crashes at symmlq.py:162 (``self.matvec`` does not exist); this
implementation does what that line means, ``y = self.op * v``.
"""
import logging

import numpy as np

from ..generic import KrylovMethod
from ..tools.utils import machine_epsilon
from .. import _engine

__docformat__ = "restructuredtext"


class Symmlq(KrylovMethod):
    """SYMMLQ for symmetric (possibly indefinite) ``(A - shift I) x = b``."""

    def __init__(self, op, **kwargs):
        KrylovMethod.__init__(self, op, **kwargs)
        self.name = "Symmetric Indefinite Lanczos with Orthogonal Factorization"
        self.acronym = "SYMMLQ"
        self.prefix = self.acronym + ": "
        self.iterates = []

    def solve(self, rhs, **kwargs):
        n = rhs.shape[0]
        nMatvec = 0
        matvec_max = kwargs.get("matvec_max", 2 * n + 2)
        rtol = kwargs.get("rtol", 1.0e-9)
        check = kwargs.get("check", False)
        shift = kwargs.get("shift", None)
        if shift == 0.0:
            shift = None
        eps = machine_epsilon()
        store_iterates = kwargs.get("store_iterates", False)
        verbose = self.logger.isEnabledFor(logging.INFO) and bool(self.logger.handlers) and \
            not all(isinstance(h, logging.NullHandler) for h in self.logger.handlers)
        first, last = "Enter SYMMLQ.   ", "Exit  SYMMLQ.   "
        msg = {-1: " beta2 = 0.  If M = I, b and x are eigenvectors",
               0: " beta1 = 0.  The exact solution is  x = 0",
               1: " Requested accuracy achieved, as determined by rtol",
               2: " Reasonable accuracy achieved, given eps",
               3: " x has converged to an eigenvector",
               4: " acond has exceeded 0.1/eps",
               5: " The iteration limit was reached",
               6: " aprod  does not define a symmetric matrix",
               7: " msolve does not define a symmetric matrix",
               8: " msolve does not define a pos-def preconditioner"}
        log = self.logger.info
        log(first + "Solution of symmetric Ax = b")
        log("n     =  %3g    precon =  %5s           " % (n, repr(self.precon is None)))
        if shift is not None:
            log("shift  =  %23.14e" % shift)
        log("maxit =  %3g     eps    =  %11.2e    rtol   =  %11.2e" % (int((matvec_max - 2.0) / 2), eps, rtol))

        op, precon = self.op, self.precon
        B = _engine.HostBridge(n, self.context, op)
        istop = itn = 0
        ynorm = acond = anorm = xnorm = rnorm = 0
        done = False
        x, w, v = B.vec(), B.vec(), B.vec()
        if store_iterates:
            self.iterates.append(x.download())
        rhs64 = np.asarray(rhs, dtype=np.float64)
        r1 = B.vec(rhs64)                                          # symmlq.py:122
        y = B.vec(rhs64) if precon is None else B.apply_precon(precon, r1, B.vec())
        r2, tmp = B.vec(), B.vec()
        b1 = y.peek(0)
        beta1 = B.fused([], [(r1, y)])[0]
        if check and precon is not None:                           # :131-139
            B.apply_precon(precon, y, r2)
            s, t = B.fused([], [(y, y), (r1, r2)])
            if np.abs(s - t) > (s + eps) * eps ** (1.0 / 3):
                istop, done = 7, True
        if beta1 < 0:
            istop, done = 8, True
        if beta1 == 0:
            done = True
        x1cg = cgnorm = qrnorm = bstep = lqnorm = 0
        diag = 1.0
        if beta1 > 0:
            beta1 = np.sqrt(beta1)
            B.fused([dict(z=v, u=y, a=1.0 / beta1)])               # v = s*y, :152-153
            B.apply(op, v, y)                                      # :162 (fixed)
            nMatvec += 1
            if check:
                B.apply(op, y, r2)
                s, t = B.fused([], [(y, y), (v, r2)])
                if abs(s - t) > (s + eps) * eps ** (1.0 / 3):
                    istop, done = 6, True
            ops = [dict(z=y, u=y, w=v, a=1.0, b=-shift)] if shift is not None else []
            alfa = B.fused(ops, [(v, y)])[0]                       # :173-174
            z, s = B.fused([dict(z=y, u=y, w=r1, a=1.0, b=-(alfa / beta1))], [(v, y), (v, v)])   # :175-180
            B.fused([dict(z=y, u=y, w=v, a=1.0, b=-(z / s)), dict(z=r2, u=y, a=1.0)])   # (r2 copies the NEW y)
            if precon is not None:
                B.apply_precon(precon, r2, y)
            oldb = beta1
            beta, r2r2, vr2 = B.fused([], [(r2, y), (r2, r2), (v, r2)])
            if beta < 0:
                istop, done = 8, True
            beta = np.sqrt(beta)
            if beta <= eps:
                istop = -1
            denom = np.sqrt(s) * np.sqrt(r2r2) + eps               # :196-199
            s = z / denom
            t = vr2 / denom
            log("beta1 =  %10.2e   alpha1 =  %9.2e" % (beta1, alfa))
            log("(v1, v2) before and after  %14.2e" % s)
            log("local reorthogonalization  %14.2e" % t)
            cgnorm, rhs2, tnorm = beta1, 0, alfa ** 2 + beta ** 2  # :206-211
            gbar, bstep, ynorm2 = alfa, 0, 0
            dbar, snprod, gmax = beta, 1, np.abs(alfa) + eps
            rhs1, x1cg, gmin = beta1, 0, np.abs(alfa) + eps
            qrnorm = beta1
        log("   Itn     x(1)(cg)  normr(cg)  r(minres)    bstep    anorm    acond")
        log("%6g %12.5e %10.3e %10.3e  %8.1e" % (itn, x1cg, cgnorm, qrnorm, (bstep / beta1) if beta1 else 0.0))

        csr = _engine.plane_csr(op)
        on_device = (csr is not None and precon is None and not verbose and not store_iterates and not done)
        if on_device:
            # device-resident loop: the scalar plane of symmlq.py:235-355 runs in csrc/lls.cu; the host
            # enqueues whole Lanczos trips and reads one status block per check interval
            from ..device import ScalarPlane as SL
            loop = _engine.PlaneLoop(B.ctx, "symmlq")
            loop.P.setup(dict(beta1=beta1, beta=beta, oldb=oldb, alfa=alfa, tnorm=tnorm, ynorm2=ynorm2, gbar=gbar,
                              dbar=dbar, rhs1=rhs1, rhs2=rhs2, snprod=snprod, bstep=bstep, gmax=gmax, gmin=gmin,
                              cgnorm=cgnorm, qrnorm=qrnorm),
                         istop=istop, itn=itn, nmatvec=nMatvec, itnlim=matvec_max, rtol=rtol, eps=eps)

            def trip():
                loop.P.step(1)                                       # norms, stopping tests; s = 1/beta
                loop.ops([dict(z=v, u=y, a_slot=SL.C0)])             # v = s*y
                if shift is None:                                    # y = A v + c1 r1, v.y and alfa: one launch
                    loop.spmv_ops(csr, v, dict(z=y, w=r1, a=1.0, b_slot=SL.C1), step=2, dot_with=v)
                else:
                    csr.spmv(v, y)
                    loop.ops([dict(z=y, u=y, w=v, a=1.0, b=-shift), dict(z=y, u=y, w=r1, a=1.0, b_slot=SL.C1)],
                             [(v, y)], step=2)                       # alfa
                loop.ops([dict(z=y, u=y, w=r2, a=1.0, b_slot=SL.C2), dict(z=r1, u=r2, a=1.0), dict(z=r2, u=y, a=1.0)],
                         [(r2, y)], step=3)                          # beta, rotation, step lengths
                loop.ops([dict(z=tmp, u=w, w=v, a_slot=SL.C3, b_slot=SL.C4), dict(z=x, u=x, w=tmp, a=1.0, b=1.0),
                          dict(z=w, u=w, w=v, a_slot=SL.C5, b_slot=SL.C6, b_neg=1)])

            st_, sc = loop.run(trip, self.check_interval)
            istop, itn = int(st_.istop), int(st_.itn)
            op._nMatvec += int(st_.nmatvec) - nMatvec
            nMatvec = int(st_.nmatvec)
            anorm, ynorm, acond = sc["anorm"], sc["ynorm"], sc["acond"]
            cgnorm, lqnorm, qrnorm, diag = sc["cgnorm"], sc["lqnorm"], sc["qrnorm"], sc["diag"]
            rhs1, snprod, bstep, ynorm2 = sc["rhs1"], sc["snprod"], sc["bstep"], sc["ynorm2"]

        if not done and not on_device:
            while nMatvec < matvec_max:
                itn += 1
                anorm = np.sqrt(tnorm)
                ynorm = np.sqrt(ynorm2)
                epsa = anorm * eps
                epsx = anorm * ynorm * eps
                epsr = anorm * ynorm * rtol
                diag = gbar
                if diag == 0:
                    diag = epsa
                lqnorm = np.sqrt(rhs1 ** 2 + rhs2 ** 2)
                qrnorm = snprod * beta1
                cgnorm = qrnorm * beta / np.abs(diag)
                if lqnorm < cgnorm:                                # :257-261
                    acond = gmax / gmin
                else:
                    acond = gmax / min(gmin, np.abs(diag))
                zbar = rhs1 / diag
                z = (snprod * zbar + bstep) / beta1
                if istop == 0:                                     # :271-276
                    if nMatvec >= matvec_max:
                        istop = 5
                    if acond >= 0.1 / eps:
                        istop = 4
                    if epsx >= beta1:
                        istop = 3
                    if cgnorm <= epsx:
                        istop = 2
                    if cgnorm <= epsr:
                        istop = 1
                if verbose and (n <= 40 or nMatvec <= 20 or nMatvec >= matvec_max - 10 or itn % 10 == 0
                                or cgnorm <= 10.0 * epsx or cgnorm <= 10.0 * epsr or acond >= 0.01 / eps
                                or istop != 0):
                    x1cg = x.peek(0) + w.peek(0) * zbar + b1 * z
                    log("%6g %12.5e %10.3e %10.3e  %8.1e %8.1e %8.1e"
                        % (itn, x1cg, cgnorm, qrnorm, bstep / beta1, anorm, acond))
                if istop != 0:
                    break
                # Lanczos step, :300-313
                B.fused([dict(z=v, u=y, a=1 / beta)])              # v = s*y
                B.apply(op, v, y)
                nMatvec += 1
                ops = [dict(z=y, u=y, w=v, a=1.0, b=-shift)] if shift is not None else []
                ops.append(dict(z=y, u=y, w=r1, a=1.0, b=-(beta / oldb)))
                alfa = B.fused(ops, [(v, y)])[0]
                ops = [dict(z=y, u=y, w=r2, a=1.0, b=-(alfa / beta)), dict(z=r1, u=r2, a=1.0),
                       dict(z=r2, u=y, a=1.0)]
                if precon is None:
                    beta_sq = B.fused(ops, [(r2, y)])[0]
                else:
                    B.fused(ops)
                    B.apply_precon(precon, r2, y)
                    beta_sq = B.fused([], [(r2, y)])[0]
                oldb = beta
                beta = beta_sq
                if beta < 0:
                    istop = 6
                    break
                beta = np.sqrt(beta)
                tnorm = tnorm + alfa ** 2 + oldb ** 2 + beta ** 2
                gamma = np.sqrt(gbar ** 2 + oldb ** 2)             # :322-328
                cs = gbar / gamma
                sn = oldb / gamma
                delta = cs * dbar + sn * alfa
                gbar = sn * dbar - cs * alfa
                epsln = sn * beta
                dbar = -cs * beta
                z = rhs1 / gamma                                   # :332-336
                s = z * cs
                t = z * sn
                B.fused([dict(z=tmp, u=w, w=v, a=s, b=t), dict(z=x, u=x, w=tmp, a=1.0, b=1.0),
                         dict(z=w, u=w, w=v, a=sn, b=-cs)])
                if store_iterates:
                    self.iterates.append(x.download())
                bstep = snprod * cs * z + bstep                    # :343-349
                snprod = snprod * sn
                gmax = max(gmax, gamma)
                gmin = min(gmin, gamma)
                ynorm2 = z ** 2 + ynorm2
                rhs1 = rhs2 - delta * z
                rhs2 = -epsln * z

        if cgnorm < lqnorm:                                        # :359-363 move to the CG point
            zbar = rhs1 / diag
            bstep = snprod * zbar + bstep
            ynorm = np.sqrt(ynorm2 + zbar ** 2)
            B.fused([dict(z=x, u=x, w=w, a=1.0, b=zbar)])
        bstep = (bstep / beta1) if beta1 else 0.0                  # :367-372 step along b
        ystep = B.vec(rhs64) if precon is None else B.apply_precon(precon, B.vec(rhs64), B.vec())
        B.fused([dict(z=x, u=x, w=ystep, a=1.0, b=bstep)])
        B.apply(op, x, y)                                          # :376-380 final residual
        nMatvec += 1
        ops = [dict(z=y, u=y, w=x, a=1.0, b=-shift)] if shift is not None else []
        ops.append(dict(z=r1, u=B.vec(rhs64), w=y, a=1.0, b=-1.0))
        rr, xx = B.fused(ops, [(r1, r1), (x, x)])
        rnorm, xnorm = np.sqrt(rr), np.sqrt(xx)
        log(last + " istop   =  %3g               itn   =   %5g" % (istop, itn))
        log(last + " anorm   =  %12.4e      acond =  %12.4e" % (anorm, acond))
        log(last + " rnorm   =  %12.4e      xnorm =  %12.4e" % (rnorm, xnorm))
        log(last + msg[istop])
        self.nMatvec = nMatvec
        self.bestSolution = self.x = x.download()
        self.solutionNorm = self.xNorm = xnorm
        self.residNorm = rnorm
        self.acond, self.anorm = acond, anorm
        self.istop, self.itn = istop, itn
        self.converged = istop in (1, 2, 3)
