"""Host-driven solver loops for operators / preconditioners the device cannot see.

When the operator is a Python closure (``LinearOperator(n, n, lambda x: ...)``) or the
preconditioner is an arbitrary object, the iteration cannot be device-resident: the
closure has to run on the host.  These loops keep every vector in HBM and run all
vector arithmetic and inner products as CUDA kernels (``kry_multi_axpy_dot``); only
the operator / preconditioner application crosses PCIe (``HostBridge.apply``), and
the scalar recurrence is the reference's, executed on the host.

Each function restates one reference loop (line numbers relative to
/root/reference/pykrylov/) in terms of fused device passes.
"""
import numpy as np

from . import _engine


def _pre(B, precon, src, dst):
    """dst = precon * src (or alias src when there is no preconditioner)."""
    return src if precon is None else B.apply_precon(precon, src, dst)


# ------------------------------------------------------------------ Bi-CGSTAB
def bicgstab(self, rhs, guess, matvec_max, result_type):
    """bicgstab/bicgstab.py:52-151."""
    n = rhs.shape[0]
    B = _engine.HostBridge(n, self.context, self.op)
    op, precon, log = self.op, self.precon, self.logger.info
    nMatvec = 0
    x = B.vec(guess)
    b = B.vec(np.asarray(rhs, dtype=np.float64))
    r0, tmp = B.vec(), B.vec()
    if guess is not None:
        B.apply(op, x, tmp)
        nMatvec += 1
        B.fused([dict(z=r0, u=b, w=tmp, a=1.0, b=-1.0)])                     # :64
    else:
        B.fused([dict(z=r0, u=b, a=1.0)])
    rho = alpha = omega = 1.0
    rho_next = B.fused([], [(r0, r0)])[0]
    residNorm = self.residNorm0 = np.abs(np.sqrt(rho_next))
    threshold = max(self.abstol, self.reltol * self.residNorm0)
    finished = (residNorm <= threshold or nMatvec >= matvec_max)
    log("Initial residual = %8.2e" % self.residNorm0)
    log("Threshold = %8.2e" % threshold)
    hdr = "%6s  %8s" % ("Matvec", "Residual")
    log(hdr)
    log("-" * len(hdr))
    r, p, v, s, t = B.vec(), B.vec(), B.vec(), B.vec(), B.vec()
    q_buf, z_buf = (B.vec(), B.vec()) if precon is not None else (None, None)
    if not finished:
        B.fused([dict(z=r, u=r0, a=1.0)])
    while not finished:
        beta = rho_next / rho * alpha / omega                                # :87
        rho = rho_next
        B.fused([dict(z=p, u=p, w=v, a=beta, b=-(beta * omega)), dict(z=p, u=p, w=r, a=1.0, b=1.0)])   # :91-93
        q = _pre(B, precon, p, q_buf)
        B.apply(op, q, v)
        nMatvec += 1
        alpha = rho / B.fused([], [(r0, v)])[0]                              # :103
        residNorm = np.sqrt(B.fused([dict(z=s, u=r, w=v, a=1.0, b=-alpha)], [(s, s)])[0])   # :104-107
        log("%6d  %8.2e" % (nMatvec, residNorm))
        if residNorm <= threshold:                                           # :110-113
            B.fused([dict(z=x, u=x, w=q, a=1.0, b=alpha)])
            finished = True
            continue
        if nMatvec >= matvec_max:
            finished = True
            continue
        z = _pre(B, precon, s, z_buf)
        B.apply(op, z, t)
        nMatvec += 1
        ts, tt, r0t = B.fused([], [(t, s), (t, t), (r0, t)])                 # :126-127
        omega = ts / tt
        rho_next = -omega * r0t
        rr = B.fused([dict(z=r, u=s, w=t, a=1.0, b=-omega), dict(z=z, u=z, a=omega),
                      dict(z=x, u=x, w=z, a=1.0, b=1.0), dict(z=x, u=x, w=q, a=1.0, b=alpha)],
                     [(r, r)])[0]                                            # :130-139
        residNorm = np.sqrt(rr)
        log("%6d  %8.2e" % (nMatvec, residNorm))
        if residNorm <= threshold or nMatvec >= matvec_max:
            finished = True
    self.converged = bool(residNorm <= threshold)
    self.nMatvec = nMatvec
    self.bestSolution = self.x = x.download().astype(result_type, copy=False)
    self.residNorm = residNorm


# ------------------------------------------------------------------------ CGS
def cgs(self, rhs, guess, matvec_max, result_type):
    """cgs/cgs.py:49-123 (the initial-residual product is not counted, :59-60)."""
    n = rhs.shape[0]
    B = _engine.HostBridge(n, self.context, self.op)
    op, precon, log = self.op, self.precon, self.logger.info
    nMatvec = 0
    x = B.vec(guess)
    b = B.vec(np.asarray(rhs, dtype=np.float64))
    r0, tmp = B.vec(), B.vec()
    if guess is not None:
        B.apply(op, x, tmp)
        B.fused([dict(z=r0, u=b, w=tmp, a=1.0, b=-1.0)])
    else:
        B.fused([dict(z=r0, u=b, a=1.0)])
    rho = B.fused([], [(r0, r0)])[0]
    residNorm = np.abs(np.sqrt(rho))
    self.residNorm0 = residNorm
    threshold = max(self.abstol, self.reltol * self.residNorm0)
    log("Initial residual = %8.2e\n" % self.residNorm0)
    log("Threshold = %8.2e\n" % threshold)
    finished = (residNorm <= threshold or nMatvec >= matvec_max)
    r, u, p, q, v, z, Az = (B.vec() for _ in range(7))
    y_buf, upq = (B.vec() if precon is not None else None), B.vec()
    if not finished:
        B.fused([dict(z=r, u=r0, a=1.0), dict(z=u, u=r0, a=1.0), dict(z=p, u=r0, a=1.0)])   # :72-74
    while not finished:
        y = _pre(B, precon, p, y_buf)
        B.apply(op, y, v)
        nMatvec += 1
        sigma = B.fused([], [(r0, v)])[0]
        alpha = rho / sigma                                                  # :86
        if precon is None:
            B.fused([dict(z=q, u=u, w=v, a=1.0, b=-alpha), dict(z=z, u=u, w=q, a=1.0, b=1.0),
                     dict(z=x, u=x, w=z, a=1.0, b=alpha)])                   # :87-95
        else:
            B.fused([dict(z=q, u=u, w=v, a=1.0, b=-alpha), dict(z=upq, u=u, w=q, a=1.0, b=1.0)])
            B.apply_precon(precon, upq, z)
            B.fused([dict(z=x, u=x, w=z, a=1.0, b=alpha)])
        B.apply(op, z, Az)
        nMatvec += 1
        rr, r0r = B.fused([dict(z=r, u=r, w=Az, a=1.0, b=-alpha)], [(r, r), (r0, r)])   # :97-106
        residNorm = np.sqrt(rr)
        if residNorm <= threshold or nMatvec >= matvec_max:
            finished = True
            continue
        rho_next = r0r
        beta = rho_next / rho
        rho = rho_next
        B.fused([dict(z=u, u=r, w=q, a=1.0, b=beta),                         # :109
                 dict(z=p, u=p, w=q, a=beta, b=1.0),                         # p*=beta; p+=q
                 dict(z=p, u=p, w=u, a=beta, b=1.0)])                        # p*=beta; p+=u
        log("%5d  %8.2e\n" % (nMatvec, residNorm))
    self.converged = bool(residNorm <= threshold)
    self.nMatvec = nMatvec
    self.bestSolution = self.x = x.download().astype(result_type, copy=False)
    self.residNorm = residNorm


# ---------------------------------------------------------------------- TFQMR
def tfqmr(self, rhs, guess, matvec_max, result_type):
    """tfqmr/tfqmr.py:48-160."""
    n = rhs.shape[0]
    B = _engine.HostBridge(n, self.context, self.op)
    op, precon, log = self.op, self.precon, self.logger.info
    nMatvec = 0
    x = B.vec(guess)
    b = B.vec(np.asarray(rhs, dtype=np.float64))
    r0, tmp = B.vec(), B.vec()
    if guess is not None:
        B.apply(op, x, tmp)
        B.fused([dict(z=r0, u=b, w=tmp, a=1.0, b=-1.0)])
    else:
        B.fused([dict(z=r0, u=b, a=1.0)])
    rho = B.fused([], [(r0, r0)])[0]
    residNorm = np.abs(np.sqrt(rho))
    self.residNorm0 = residNorm
    threshold = max(self.abstol, self.reltol * self.residNorm0)
    log("Initial residual = %8.2e" % self.residNorm0)
    log("Threshold = %8.2e" % threshold)
    finished = (residNorm <= threshold or nMatvec >= matvec_max)
    m = None
    y, w, d, u, v = (B.vec() for _ in range(5))
    z_buf = B.vec() if precon is not None else None
    if not finished:
        B.fused([dict(z=y, u=r0, a=1.0), dict(z=w, u=r0, a=1.0)])            # :71-72 (d = 0 already)
        theta = eta = 0.0
        k = 0
        z = _pre(B, precon, y, z_buf)
        B.apply(op, z, u)
        nMatvec += 1
        B.fused([dict(z=v, u=u, a=1.0)])

    def half(alpha, theta, eta, residNorm):
        """w -= alpha u ; d = (theta^2 eta/alpha) d + z ; theta, c, tau, eta ; x += eta d  (:92-99)."""
        coef = theta * theta * eta / alpha
        ww = B.fused([dict(z=w, u=w, w=u, a=1.0, b=-alpha), dict(z=d, u=d, w=z, a=coef, b=1.0)], [(w, w)])[0]
        theta = np.sqrt(ww) / residNorm
        c = 1.0 / np.sqrt(1 + theta * theta)
        residNorm *= theta * c
        eta = c * c * alpha
        B.fused([dict(z=x, u=x, w=d, a=1.0, b=eta)])
        return theta, eta, residNorm

    while not finished:
        k += 1
        sigma = B.fused([], [(r0, v)])[0]
        alpha = rho / sigma
        theta, eta, residNorm = half(alpha, theta, eta, residNorm)           # first pass
        m = 2.0 * k - 1.0
        if residNorm * np.sqrt(m + 1) < threshold or nMatvec >= matvec_max:
            finished = True
            continue
        m += 1                                                               # second pass
        B.fused([dict(z=y, u=y, w=v, a=1.0, b=-alpha)])                      # :109
        z = _pre(B, precon, y, z_buf)
        B.apply(op, z, u)
        nMatvec += 1
        theta, eta, residNorm = half(alpha, theta, eta, residNorm)
        if residNorm * np.sqrt(m + 1) < threshold or nMatvec >= matvec_max:
            finished = True
            continue
        rho_next = B.fused([], [(r0, w)])[0]                                 # :128
        beta = rho_next / rho
        rho = rho_next
        B.fused([dict(z=y, u=y, w=w, a=beta, b=1.0),                         # :133-134
                 dict(z=v, u=v, w=u, a=beta, b=1.0), dict(z=v, u=v, a=beta)])   # :137-139
        z = _pre(B, precon, y, z_buf)
        B.apply(op, z, u)
        nMatvec += 1
        B.fused([dict(z=v, u=v, w=u, a=1.0, b=1.0)])                         # :149
        log("%5d  %8.2e" % (nMatvec, residNorm))
    if m is None:
        self.converged = False       # the reference raises NameError here (tfqmr.py:156)
    else:
        self.converged = bool(residNorm * np.sqrt(m + 1) < threshold)
    self.nMatvec = nMatvec
    self.bestSolution = self.x = x.download().astype(result_type, copy=False)
    self.residNorm = residNorm


# --------------------------------------------------------------------- MINRES
def minres(self, b, precon, shift, show, check, itnlim, rtol, etol, store_iterates, window, result_type):
    """minres/minres.py:132-408 with the vectors in HBM (opaque operator and/or
    preconditioner).  Returns nothing; fills the attributes like the reference."""
    from math import sqrt
    from .tools.utils import check_symmetric
    A = self.op
    n = b.shape[0]
    eps = self.eps
    B = _engine.HostBridge(n, self.context, A)
    x = B.vec()
    xNrgNorm2 = 0.0
    dErr = np.zeros(window)
    trncDirErr = 0
    if store_iterates:
        self.iterates.append(x.download())
    istop = itn = 0
    Anorm = Acond = rnorm = ynorm = 0.0
    done = False
    b64 = np.asarray(b, dtype=np.float64)
    r1 = B.vec(b64)
    y = B.vec(b64) if precon is None else B.apply_precon(precon, r1, B.vec())
    beta1 = B.fused([], [(r1, y)])[0]
    if beta1 < 0:
        istop, done = 9, True
    if beta1 == 0.0:
        done = True
    if beta1 > 0:
        beta1 = sqrt(beta1)
    self.residNorm0 = beta1
    if check:
        if not check_symmetric(A):
            istop, done = 7, True
    if check and (precon is not None):
        if not check_symmetric(precon):
            istop, done = 8, True
    oldb, beta, dbar, epsln = 0.0, beta1, 0.0, 0.0
    qrnorm = phibar = rhs1 = beta1
    Arnorm = rhs2 = tnorm2 = ynorm2 = 0.0
    cs, sn = -1.0, 0.0
    w, w2, w1, v, r2 = B.vec(), B.vec(), B.vec(), B.vec(), B.vec(b64)
    gmax = gmin = 0.0
    if show:
        print(" " * 2)
        print("   Itn     x[0]     Compatible    LS" + "       norm(A)  cond(A) gbar/|A|")
    if not done:
        while itn < itnlim:
            itn += 1
            s = 1.0 / beta
            B.fused([dict(z=v, u=y, a=s)])                                   # :237
            B.apply(A, v, y)                                                 # :239
            ops = [dict(z=y, u=y, w=v, a=1.0, b=-shift)]                     # :240
            if itn >= 2:
                ops.append(dict(z=y, u=y, w=r1, a=1.0, b=-(beta / oldb)))    # :243
            alfa = B.fused(ops, [(v, y)])[0]                                 # :245
            ops = [dict(z=y, u=r2, w=y, a=(-alfa / beta), b=1.0),            # :246
                   dict(z=r1, u=r2, a=1.0), dict(z=r2, u=y, a=1.0)]          # :247-248
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
            beta = sqrt(beta)
            tnorm2 = tnorm2 + alfa ** 2 + oldb ** 2 + beta ** 2
            if itn == 1:
                if beta / beta1 <= 10 * eps:
                    istop = -1
                gmax = abs(alfa)
                gmin = gmax
            oldeps = epsln
            delta = cs * dbar + sn * alfa
            gbar = sn * dbar - cs * alfa
            epsln = sn * beta
            dbar = -cs * beta
            root = self.normof2(gbar, dbar)
            Arnorm = phibar * root
            gamma = self.normof2(gbar, beta)
            gamma = max(gamma, eps)
            cs = gbar / gamma
            sn = beta / gamma
            phi = cs * phibar
            phibar = sn * phibar
            denom = 1.0 / gamma
            # w1 = w2 ; w2 = w ; w = (v - oldeps w1 - delta w2) denom ; x += phi w    (:294-297)
            B.fused([dict(z=w1, u=w2, a=1.0), dict(z=w2, u=w, a=1.0),
                     dict(z=w, u=v, w=w1, a=1.0, b=-oldeps), dict(z=w, u=w, w=w2, a=1.0, b=-delta)])
            B.fused([dict(z=w, u=w, a=denom), dict(z=x, u=x, w=w, a=1.0, b=phi)])
            if store_iterates:
                self.iterates.append(x.download())
            xNrgNorm2 += phi * phi
            dErr[itn % window] = phi
            if itn > window:
                trncDirErr = np.linalg.norm(dErr)
                xNrgNorm = sqrt(xNrgNorm2)
                self.dir_errors_window.append(trncDirErr / xNrgNorm)
                if trncDirErr < etol * xNrgNorm:
                    istop = 10
            gmax = max(gmax, gamma)
            gmin = min(gmin, gamma)
            z = rhs1 / gamma
            ynorm2 = z ** 2 + ynorm2
            rhs1 = rhs2 - delta * z
            rhs2 = -epsln * z
            Anorm = sqrt(tnorm2)
            ynorm = sqrt(ynorm2)
            epsa = Anorm * eps
            epsx = Anorm * ynorm * eps
            epsr = Anorm * ynorm * rtol
            diag = gbar
            if diag == 0:
                diag = epsa
            qrnorm = phibar
            rnorm = qrnorm
            test1 = rnorm / (Anorm * ynorm)
            test2 = root / Anorm
            self.residHistory.append(rnorm)
            Acond = gmax / gmin
            if istop == 0:
                t1 = 1 + test1
                t2 = 1 + test2
                if t2 <= 1:
                    istop = 2
                if t1 <= 1:
                    istop = 1
                if itn >= itnlim:
                    istop = 6
                if Acond >= 0.1 / eps:
                    istop = 4
                if epsx >= beta1:
                    istop = 3
                if test2 <= rtol:
                    istop = 2
                if test1 <= rtol:
                    istop = 1
            prnt = (n <= 40 or itn <= 10 or itn >= itnlim - 10 or (itn % 10) == 0 or qrnorm <= 10 * epsx
                    or qrnorm <= 10 * epsr or Acond <= 1e-2 / eps or istop != 0)
            if show and prnt:
                print("%6g %12.5e %10.3e %10.3e %8.1e %8.1e %8.1e"
                      % (itn, x.peek(0), test1, test2, Anorm, Acond, gbar / Anorm))
            if istop > 0:
                break
            if (itn % 10) == 0:
                print(" ")                                 # (sic) minres.py:383 ignores `show`
    if show:
        last = self.last
        print(last + " istop   =  %3g               itn   =%5g" % (istop, itn))
        print(last + " Anorm   =  %12.4e      Acond =  %12.4e" % (Anorm, Acond))
        print(last + " rnorm   =  %12.4e      ynorm =  %12.4e" % (rnorm, ynorm))
        print(last + " Arnorm  =  %12.4e" % Arnorm)
        print(last + self.msg[istop + 1])
    self.converged = istop in [1, 2, 3, 4, 10]
    if istop == 10:
        self.status = "direct error small"
    self.x = self.bestSolution = x.download().astype(result_type, copy=False)
    self.istop = istop
    self.itn = self.nMatvec = itn
    self.rnorm = self.residNorm = rnorm
    self.Arnorm = Arnorm
    self.Anorm = Anorm
    self.Acond = Acond
    self.ynorm = ynorm
