"""CPU restatement of the reference's Krylov hot path (NumPy, fp64).

TEST INFRASTRUCTURE ONLY -- this module is the *checker*.  It is imported by
tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl
reference legs, and by nothing else.  The product package (pykrylov_b200) never
imports it and has no CPU fallback.

Every function restates one reference loop *as a resumable state machine*
(``*_start`` builds the state, ``*_step`` advances one iteration) so the tests
can do single-step parity from an identical state.
The floating-point operation order of each update is the reference's, so on
the same NumPy the port is bit-identical to the reference itself; that is
pinned by tests/test_oracle.py against (a) the transliterated reference in
oracle/_ref when it is present and (b) the committed vectors in tests/golden/
which were produced by the reference (tests/golden/make_golden.py).

Reference citations are relative to /root/reference/pykrylov/.
"""
from math import sqrt

import numpy as np

EPS = np.float64(np.finfo(np.double).eps)          # tools/utils.py:7-9


class State(dict):
    """dict with attribute access; holds one solver's vectors and scalars."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def _apply(M, v):
    """Apply an optional preconditioner / operator callable."""
    return v if M is None else M(v)


# --------------------------------------------------------------------------
# CG                                                        cg/cg.py:50-165
# --------------------------------------------------------------------------
def cg_start(A, rhs, guess=None, precon=None, abstol=1.0e-8, reltol=1.0e-6,
             matvec_max=None, check_curvature=True):
    n = rhs.shape[0]
    st = State(kind="cg", n=n, nMatvec=0, definite=True, precon=precon,
               check_curvature=check_curvature, infiniteDescent=None)
    st.x = (np.zeros(n) if guess is None else guess).astype(np.float64)   # cg.py:77
    st.matvec_max = 2 * n if matvec_max is None else matvec_max             # cg.py:82
    r = -rhs                                                                # cg.py:85
    if guess is not None:
        r += A(st.x)                                                        # cg.py:87-88
        st.nMatvec += 1
    st.r = r
    y = _apply(precon, r)                                                   # cg.py:91-94
    st.ry = np.float64(np.dot(r, y))                                             # cg.py:99
    st.residNorm0 = st.residNorm = np.float64(np.abs(np.sqrt(st.ry)))            # cg.py:100
    st.residHistory = [st.residNorm0]
    st.threshold = max(abstol, reltol * st.residNorm0)                      # cg.py:102
    st.p = -r                                                               # cg.py:104
    st.pAp = float("nan")
    st.alpha = st.beta = float("nan")
    return st


def cg_active(st):
    return st.residNorm > st.threshold and st.nMatvec < st.matvec_max and st.definite  # cg.py:113


def cg_step(A, st):
    """One pass of the body of cg.py:113-158."""
    Ap = A(st.p)                                                            # cg.py:115
    st.nMatvec += 1
    st.Ap = Ap
    pAp = np.float64(np.dot(st.p, Ap))                                           # cg.py:117
    st.pAp = pAp
    if st.check_curvature and pAp <= 0:                                     # cg.py:119-124
        st.infiniteDescent = st.p
        st.definite = False
        return st
    alpha = st.ry / pAp                                                     # cg.py:127
    st.alpha = alpha
    st.x += alpha * st.p                                                    # cg.py:130
    st.r += alpha * Ap                                                      # cg.py:131
    y = _apply(st.precon, st.r)                                             # cg.py:137-140
    ry_next = np.float64(np.dot(st.r, y))                                        # cg.py:146
    beta = ry_next / st.ry                                                  # cg.py:149
    st.beta = beta
    st.p *= beta                                                            # cg.py:150
    st.p -= st.r                                                            # cg.py:151
    st.ry = ry_next
    st.residNorm = np.float64(np.abs(np.sqrt(st.ry)))                            # cg.py:154
    st.residHistory.append(st.residNorm)
    return st


def cg_solve(A, rhs, **kw):
    st = cg_start(A, rhs, **kw)
    while cg_active(st):
        cg_step(A, st)
    st.converged = st.residNorm <= st.threshold                             # cg.py:161
    return st


# --------------------------------------------------------------------------
# Bi-CGSTAB                                      bicgstab/bicgstab.py:43-151
# --------------------------------------------------------------------------
def bicgstab_start(A, rhs, guess=None, precon=None, abstol=1.0e-8, reltol=1.0e-6,
                   matvec_max=None):
    n = rhs.shape[0]
    st = State(kind="bicgstab", n=n, nMatvec=0, precon=precon)
    st.x = (np.zeros(n) if guess is None else guess).astype(np.float64)
    st.matvec_max = 2 * n if matvec_max is None else matvec_max
    r0 = rhs
    if guess is not None:
        r0 = rhs - A(st.x)                                                  # bicgstab.py:64
        st.nMatvec += 1
    st.r0 = r0
    st.rho = st.alpha = st.omega = 1.0                                      # bicgstab.py:67
    st.rho_next = np.float64(np.dot(r0, r0))                                     # bicgstab.py:68
    st.residNorm = st.residNorm0 = np.float64(np.abs(np.sqrt(st.rho_next)))
    st.threshold = max(abstol, reltol * st.residNorm0)
    st.finished = bool(st.residNorm <= st.threshold or st.nMatvec >= st.matvec_max)
    st.residHistory = [st.residNorm0]
    if not st.finished:
        st.r = r0.copy()                                                    # bicgstab.py:81-83
        st.p = np.zeros(n)
        st.v = np.zeros(n)
    return st


def bicgstab_step(A, st):
    """One pass of the body of bicgstab.py:85-145."""
    beta = st.rho_next / st.rho * st.alpha / st.omega                       # :87
    st.beta = beta
    st.rho = st.rho_next
    p = st.p
    p *= beta                                                               # :91
    p -= beta * st.omega * st.v                                             # :92
    p += st.r                                                               # :93
    q = _apply(st.precon, p)
    st.v = A(q)                                                             # :101
    st.nMatvec += 1
    st.r0v = np.float64(np.dot(st.r0, st.v))
    st.alpha = st.rho / st.r0v                                              # :103
    s = st.r - st.alpha * st.v                                              # :104
    st.s = s
    st.residNorm = np.float64(np.linalg.norm(s))                                 # :107
    st.residHistory.append(st.residNorm)
    if st.residNorm <= st.threshold:                                        # :110-113
        st.x += st.alpha * q
        st.finished = True
        return st
    if st.nMatvec >= st.matvec_max:                                         # :115-117
        st.finished = True
        return st
    z = _apply(st.precon, s)
    t = A(z)                                                                # :125
    st.nMatvec += 1
    st.t = t
    st.ts, st.tt, st.r0t = np.float64(np.dot(t, s)), np.float64(np.dot(t, t)), np.float64(np.dot(st.r0, t))
    st.omega = st.ts / st.tt                                                # :126
    st.rho_next = -st.omega * st.r0t                                        # :127
    st.r = s - st.omega * t                                                 # :130
    z *= st.omega                                                           # :135 (aliases s)
    st.x += z                                                               # :136
    st.x += st.alpha * q                                                    # :137
    st.residNorm = np.float64(np.linalg.norm(st.r))                              # :139
    st.residHistory.append(st.residNorm)
    if st.residNorm <= st.threshold or st.nMatvec >= st.matvec_max:         # :142
        st.finished = True
    return st


def bicgstab_solve(A, rhs, **kw):
    st = bicgstab_start(A, rhs, **kw)
    while not st.finished:
        bicgstab_step(A, st)
    st.converged = st.residNorm <= st.threshold
    return st


# --------------------------------------------------------------------------
# CGS                                                      cgs/cgs.py:41-123
# --------------------------------------------------------------------------
def cgs_start(A, rhs, guess=None, precon=None, abstol=1.0e-8, reltol=1.0e-6,
              matvec_max=None):
    n = rhs.shape[0]
    st = State(kind="cgs", n=n, nMatvec=0, precon=precon)
    st.x = (np.zeros(n) if guess is None else guess).astype(np.float64)
    st.matvec_max = 2 * n if matvec_max is None else matvec_max
    r0 = rhs
    if guess is not None:
        r0 = rhs - A(st.x)                       # cgs.py:59-60 -- NOT counted in nMatvec
    st.r0 = r0
    st.rho = np.float64(np.dot(r0, r0))                                          # :62
    st.residNorm = st.residNorm0 = np.float64(np.abs(np.sqrt(st.rho)))
    st.threshold = max(abstol, reltol * st.residNorm0)
    st.finished = bool(st.residNorm <= st.threshold or st.nMatvec >= st.matvec_max)
    st.residHistory = [st.residNorm0]
    if not st.finished:
        st.r = r0.copy()                                                    # :72-74
        st.u = r0
        st.p = r0.copy()
    return st


def cgs_step(A, st):
    """One pass of the body of cgs.py:76-117."""
    y = _apply(st.precon, st.p)
    v = A(y)                                                                # :84
    st.nMatvec += 1
    st.sigma = np.float64(np.dot(st.r0, v))                                      # :85
    alpha = st.rho / st.sigma                                               # :86
    st.alpha = alpha
    q = st.u - alpha * v                                                    # :87
    z = _apply(st.precon, st.u + q)                                         # :89-92
    st.x += alpha * z                                                       # :95
    Az = A(z)                                                               # :96
    st.nMatvec += 1
    st.r -= alpha * Az                                                      # :97
    st.residNorm = np.float64(np.linalg.norm(st.r))                              # :100
    st.residHistory.append(st.residNorm)
    if st.residNorm <= st.threshold or st.nMatvec >= st.matvec_max:         # :102-104
        st.finished = True
        return st
    rho_next = np.float64(np.dot(st.r0, st.r))                                   # :106
    beta = rho_next / st.rho
    st.beta = beta
    st.rho = rho_next
    st.u = st.r + beta * q                                                  # :109
    p = st.p
    p *= beta                                                               # :112-115
    p += q
    p *= beta
    p += st.u
    return st


def cgs_solve(A, rhs, **kw):
    st = cgs_start(A, rhs, **kw)
    while not st.finished:
        cgs_step(A, st)
    st.converged = st.residNorm <= st.threshold
    return st


# --------------------------------------------------------------------------
# TFQMR                                                tfqmr/tfqmr.py:40-160
# --------------------------------------------------------------------------
def tfqmr_start(A, rhs, guess=None, precon=None, abstol=1.0e-8, reltol=1.0e-6,
                matvec_max=None):
    n = rhs.shape[0]
    st = State(kind="tfqmr", n=n, nMatvec=0, precon=precon)
    st.x = (np.zeros(n) if guess is None else guess).astype(np.float64)
    st.matvec_max = 2 * n if matvec_max is None else matvec_max
    r0 = rhs
    if guess is not None:
        r0 = rhs - A(st.x)                       # tfqmr.py:58-59 -- NOT counted
    st.r0 = r0
    st.rho = np.float64(np.dot(r0, r0))                                          # :61
    st.residNorm = st.residNorm0 = np.float64(np.abs(np.sqrt(st.rho)))
    st.threshold = max(abstol, reltol * st.residNorm0)
    st.finished = bool(st.residNorm <= st.threshold or st.nMatvec >= st.matvec_max)
    st.residHistory = [st.residNorm0]
    st.m = None
    if not st.finished:
        st.y = r0.copy()                                                    # :71-85
        st.w = r0.copy()
        st.d = np.zeros(n)
        st.theta = 0.0
        st.eta = 0.0
        st.k = 0
        st.z = _apply(precon, st.y)
        st.u = A(st.z)
        st.nMatvec += 1
        st.v = st.u.copy()
    return st


def _tfqmr_half(st, alpha):
    """The block repeated at tfqmr.py:92-99 and :116-123."""
    st.w -= alpha * st.u
    st.d *= st.theta * st.theta * st.eta / alpha
    st.d += st.z
    st.theta = np.float64(np.linalg.norm(st.w)) / st.residNorm
    c = 1.0 / np.sqrt(1 + st.theta * st.theta)
    st.residNorm *= st.theta * c
    st.eta = c * c * alpha
    st.x += st.eta * st.d


def tfqmr_step(A, st):
    """One pass of the body of tfqmr.py:85-153."""
    st.k += 1
    st.sigma = np.float64(np.dot(st.r0, st.v))                                   # :88
    alpha = st.rho / st.sigma
    st.alpha = alpha
    _tfqmr_half(st, alpha)                                                  # :92-99
    st.m = 2.0 * st.k - 1.0
    st.residHistory.append(np.float64(st.residNorm))
    if st.residNorm * np.sqrt(st.m + 1) < st.threshold or st.nMatvec >= st.matvec_max:
        st.finished = True
        return st
    st.m += 1                                                               # :108
    st.y -= alpha * st.v                                                    # :109
    st.z = _apply(st.precon, st.y)
    st.u = A(st.z)                                                          # :114
    st.nMatvec += 1
    _tfqmr_half(st, alpha)                                                  # :116-123
    st.residHistory.append(np.float64(st.residNorm))
    if st.residNorm * np.sqrt(st.m + 1) < st.threshold or st.nMatvec >= st.matvec_max:
        st.finished = True
        return st
    rho_next = np.float64(np.dot(st.r0, st.w))                                   # :128
    beta = rho_next / st.rho
    st.beta = beta
    st.rho = rho_next
    st.y *= beta                                                            # :133-134
    st.y += st.w
    st.v *= beta                                                            # :137-139
    st.v += st.u
    st.v *= beta
    st.z = _apply(st.precon, st.y)
    st.u = A(st.z)                                                          # :146
    st.nMatvec += 1
    st.v += st.u                                                            # :149
    return st


def tfqmr_solve(A, rhs, **kw):
    st = tfqmr_start(A, rhs, **kw)
    while not st.finished:
        tfqmr_step(A, st)
    # tfqmr.py:156 -- NameError in the reference when converged at entry (m unset)
    st.converged = bool(st.m is not None and st.residNorm * np.sqrt(st.m + 1) < st.threshold)
    return st


# --------------------------------------------------------------------------
# MINRES                                            minres/minres.py:114-408
# --------------------------------------------------------------------------
def minres_start(A, b, precon=None, shift=0.0, itnlim=None, rtol=1.0e-12,
                 etol=1.0e-6, window=5):
    """Pre-loop part, minres.py:116-209 (the symmetry check of :186-196 is the
    caller's business -- it is 20 extra operator products, not solver state)."""
    n = b.shape[0]
    st = State(kind="minres", n=n, precon=precon, shift=shift, rtol=rtol, etol=etol,
               window=window, itnlim=5 * n if itnlim is None else itnlim)
    st.x = np.zeros(n)
    st.xNrgNorm2 = 0.0
    st.dErr = np.zeros(window)
    st.trncDirErr = 0
    st.dir_errors_window = []
    st.istop = 0
    st.itn = 0
    st.Anorm = st.Acond = st.rnorm = st.ynorm = 0.0
    st.done = False
    st.r1 = b
    st.y = b.copy() if precon is None else precon(b)                        # :161-165
    beta1 = np.float64(np.dot(b, st.y))                                          # :166
    if beta1 < 0:                                                           # :170-173
        st.istop = 9
        st.done = True
    if beta1 == 0.0:
        st.done = True
    if beta1 > 0:
        beta1 = sqrt(beta1)                                                 # :179-180
    st.beta1 = st.residNorm0 = beta1
    st.oldb = 0.0                                                           # :201-205
    st.beta = beta1
    st.dbar = st.epsln = 0.0
    st.qrnorm = st.phibar = st.rhs1 = beta1
    st.Arnorm = st.rhs2 = st.tnorm2 = st.ynorm2 = 0.0
    st.cs, st.sn = -1.0, 0.0
    st.w = np.zeros(n)
    st.w2 = np.zeros(n)
    st.r2 = st.r1.copy()
    st.gmax = st.gmin = 0.0
    st.residHistory = []
    return st


def minres_active(st):
    return (not st.done) and st.itn < st.itnlim and st.istop <= 0 and not st.get("broke", False)


def minres_step(A, st):
    """One pass of the loop body, minres.py:218-383."""
    eps = EPS
    st.itn += 1
    itn = st.itn
    s = 1.0 / st.beta                                                       # :236
    v = s * st.y                                                            # :237
    st.v = v
    y = A(v)                                                                # :239
    y -= st.shift * v                                                       # :240
    if itn >= 2:
        y = y - (st.beta / st.oldb) * st.r1                                 # :243
    alfa = np.float64(np.dot(v, y))                                              # :245
    st.alfa = alfa
    y = (-alfa / st.beta) * st.r2 + y                                       # :246
    st.r1 = st.r2.copy()                                                    # :247-248
    st.r2 = y.copy()
    if st.precon is not None:
        y = st.precon(st.r2)
    st.y = y
    st.oldb = st.beta                                                       # :250
    beta = np.float64(np.dot(st.r2, y))                                          # :251
    if beta < 0:                                                            # :252-254
        st.istop = 6
        st.broke = True
        return st
    beta = sqrt(beta)
    st.beta = beta
    st.tnorm2 = st.tnorm2 + alfa ** 2 + st.oldb ** 2 + beta ** 2            # :256
    if itn == 1:                                                            # :258-264
        if beta / st.beta1 <= 10 * eps:
            st.istop = -1
        st.gmax = abs(alfa)
        st.gmin = st.gmax
    oldeps = st.epsln                                                       # :270-272
    delta = st.cs * st.dbar + st.sn * alfa
    gbar = st.sn * st.dbar - st.cs * alfa
    st.epsln = st.sn * beta                                                 # :277-280
    st.dbar = -st.cs * beta
    root = sqrt(gbar ** 2 + st.dbar ** 2)
    st.Arnorm = st.phibar * root
    gamma = sqrt(gbar ** 2 + beta ** 2)                                     # :284-289
    gamma = max(gamma, eps)
    st.cs = gbar / gamma
    st.sn = beta / gamma
    phi = st.cs * st.phibar
    st.phibar = st.sn * st.phibar
    st.phi, st.gamma, st.delta, st.gbar = phi, gamma, delta, gbar
    denom = 1.0 / gamma                                                     # :293-297
    w1 = st.w2.copy()
    st.w2 = st.w.copy()
    st.w = (v - oldeps * w1 - delta * st.w2) * denom
    st.x += phi * st.w
    st.xNrgNorm2 += phi * phi                                               # :303-310
    st.dErr[itn % st.window] = phi
    if itn > st.window:
        st.trncDirErr = np.float64(np.linalg.norm(st.dErr))
        xNrgNorm = sqrt(st.xNrgNorm2)
        st.dir_errors_window.append(st.trncDirErr / xNrgNorm)
        if st.trncDirErr < st.etol * xNrgNorm:
            st.istop = 10
    st.gmax = max(st.gmax, gamma)                                           # :314-319
    st.gmin = min(st.gmin, gamma)
    z = st.rhs1 / gamma
    st.ynorm2 = z ** 2 + st.ynorm2
    st.rhs1 = st.rhs2 - delta * z
    st.rhs2 = -st.epsln * z
    st.Anorm = sqrt(st.tnorm2)                                              # :323-334
    st.ynorm = sqrt(st.ynorm2)
    epsx = st.Anorm * st.ynorm * eps
    st.qrnorm = st.phibar
    st.rnorm = st.qrnorm
    test1 = st.rnorm / (st.Anorm * st.ynorm)
    test2 = root / st.Anorm
    st.test1, st.test2 = test1, test2
    st.residHistory.append(st.rnorm)                                        # :336
    st.Acond = st.gmax / st.gmin                                            # :344
    if st.istop == 0:                                                       # :349-361
        t1 = 1 + test1
        t2 = 1 + test2
        if t2 <= 1:
            st.istop = 2
        if t1 <= 1:
            st.istop = 1
        if itn >= st.itnlim:
            st.istop = 6
        if st.Acond >= 0.1 / eps:
            st.istop = 4
        if epsx >= st.beta1:
            st.istop = 3
        if test2 <= st.rtol:
            st.istop = 2
        if test1 <= st.rtol:
            st.istop = 1
    return st


def minres_solve(A, b, **kw):
    st = minres_start(A, b, **kw)
    while minres_active(st):
        minres_step(A, st)
    st.converged = st.istop in [1, 2, 3, 4, 10]                             # :395
    st.nMatvec = st.itn                                                     # :403
    st.residNorm = st.rnorm
    return st


# --------------------------------------------------------------------------
# Problem definitions (gallery/gallery.py:3-29 restated as CSR generators)
# --------------------------------------------------------------------------
def poisson1d_matvec(x):
    """gallery.py:3-8."""
    y = 2 * x
    y[:-1] -= x[1:]
    y[1:] -= x[:-1]
    return y


def poisson2d_matvec(x):
    """gallery.py:10-29 (vectorised over the interior blocks; identical result:
    each y entry receives the same subtractions in the same order)."""
    n = int(sqrt(x.shape[0]))
    y = 4 * x
    y[n:] -= x[:-n]
    y[:-n] -= x[n:]
    X = x.reshape(n, n)
    Y = y.reshape(n, n)
    # blocks 0..n-2 subtract the right neighbour first (gallery.py:18-25) ...
    Y[:-1, :-1] -= X[:-1, 1:]
    Y[:-1, 1:] -= X[:-1, :-1]
    # ... the last block subtracts the left neighbour first (gallery.py:27-28)
    Y[-1, 1:] -= X[-1, :-1]
    Y[-1, :-1] -= X[-1, 1:]
    return y


def poisson2d_csr(g, row_begin=0, row_end=None):
    """5-point Laplacian of gallery.Poisson2dMatvec on a g x g grid as CSR rows
    [row_begin, row_end): diag 4, off-diagonals -1, Dirichlet, lexicographic,
    sorted int32 columns (what scipy.sparse would give)."""
    n = g * g
    row_end = n if row_end is None else row_end
    i = np.arange(row_begin, row_end, dtype=np.int64)
    col = i % g
    has = np.stack([i >= g, col > 0, np.ones_like(i, bool), col < g - 1, i < n - g], axis=1)
    cols = np.stack([i - g, i - 1, i, i + 1, i + g], axis=1)
    vals = np.broadcast_to(np.array([-1.0, -1.0, 4.0, -1.0, -1.0]), cols.shape)
    indptr = np.zeros(len(i) + 1, dtype=np.int64)
    np.cumsum(has.sum(axis=1), out=indptr[1:])
    return indptr.astype(np.int32 if indptr[-1] < 2 ** 31 else np.int64), \
        cols[has].astype(np.int32), np.ascontiguousarray(vals[has])


def convdiff3d_csr(m, gamma=0.5, row_begin=0, row_end=None):
    """7-point convection-diffusion operator of SURVEY.md section 8d config 4 on
    an m^3 grid: -Laplace (diag 6, off -1) + first-order upwind convection
    gamma per axis (diag += 3*gamma, upstream neighbour -1-gamma, downstream -1)."""
    n = m ** 3
    row_end = n if row_end is None else row_end
    i = np.arange(row_begin, row_end, dtype=np.int64)
    ix, iy, iz = i % m, (i // m) % m, i // (m * m)
    has = np.stack([iz > 0, iy > 0, ix > 0, np.ones_like(i, bool),
                    ix < m - 1, iy < m - 1, iz < m - 1], axis=1)
    cols = np.stack([i - m * m, i - m, i - 1, i, i + 1, i + m, i + m * m], axis=1)
    up = -1.0 - gamma
    vals = np.broadcast_to(np.array([up, up, up, 6.0 + 3 * gamma, -1.0, -1.0, -1.0]), cols.shape)
    indptr = np.zeros(len(i) + 1, dtype=np.int64)
    np.cumsum(has.sum(axis=1), out=indptr[1:])
    return indptr.astype(np.int32 if indptr[-1] < 2 ** 31 else np.int64), \
        cols[has].astype(np.int32), np.ascontiguousarray(vals[has])


def check_symmetric(A, n, repeats=10):
    """tools/utils.py:63-85 (same seed, same test)."""
    np.random.seed(1)
    for _ in range(repeats):
        x = np.random.random(n)
        w = A(x)
        r = A(w)
        s = np.dot(w, w)
        t = np.dot(x, r)
        z = abs(s - t)
        epsa = (s + EPS) * EPS ** (1.0 / 3)
        if z > epsa:
            return False
    return True
