"""Host-side helpers around linear operators (reference: pykrylov/tools/utils.py).

These are not on the device hot path; ``check_symmetric`` issues 2*repeats operator
products through ``op * x`` (so with a device operator they run as CUDA SpMVs).
"""
from math import copysign, sqrt

import numpy as np


def machine_epsilon():
    """Unit round-off of IEEE double (utils.py:7-9)."""
    return np.finfo(np.double).eps


def check_symmetric(op, repeats=10):
    """Randomised symmetry test, same seed and criterion as utils.py:63-85:
    for random x, compare (Ax).(Ax) with x.(A(Ax))."""
    nrow, ncol = op.shape
    if nrow != ncol:
        return False
    eps = machine_epsilon()
    # Row-sharded device operator (SPMD: local slice in, local slice out): the two inner
    # products are partial sums that only mean something after a sum over the ranks -- and
    # every rank must reach the same verdict, or some would iterate while others do not.
    csr = getattr(op, "device_csr", None)
    reduce = csr.ctx.allreduce if (csr is not None and getattr(csr, "sharded", False)) else None
    np.random.seed(1)
    for _ in range(repeats):
        x = np.random.random(ncol)
        w = op * x
        s = np.dot(w, w)
        t = np.dot(x, op * w)
        if reduce is not None:
            s, t = (float(v) for v in reduce([s, t]))
        if abs(s - t) > (s + eps) * eps ** (1.0 / 3):
            return False
    return True


def check_positive_definite(op, repeats=10, semi=False):
    """Randomised definiteness test (utils.py:88-114): v.(Av) > 0 (>= 0 if semi)."""
    nrow, ncol = op.shape
    if nrow != ncol:
        return False
    eps = machine_epsilon()
    for _ in range(repeats):
        v = np.random.random(ncol)
        vw = np.dot(v, op * v)
        if np.imag(vw) > np.sqrt(eps) * np.abs(vw):
            return False
        vw = np.real(vw)
        if (vw < 0) if semi else (vw <= 0):
            return False
    return True


def roots_quadratic(q2, q1, q0, tol=1.0e-8, nitref=1):
    """Real roots of q2 x^2 + q1 x + q0 (after GALAHAD; utils.py:12-60), with
    `nitref` Newton polishing steps."""
    a2, a1, a0 = float(q2), float(q1), float(q0)
    if a2 == 0.0:
        if a1 == 0.0:
            return [0.0] if a0 == 0.0 else []
        roots = [-a0 / a1]
    elif abs(a0 * a2) > tol * a1 * a1:
        disc = a1 * a1 - 4.0 * a2 * a0
        if disc < 0.0:
            return []
        d = -0.5 * (a1 + copysign(sqrt(disc), a1))
        roots = [d / a2, a0 / d]
    else:
        roots = [-a1 / a2, 0.0]       # ill-conditioned: one root is (numerically) zero
    polished = []
    for root in roots:
        for _ in range(nitref):
            der = 2.0 * a2 * root + a1
            if der != 0.0:
                root -= ((a2 * root + a1) * root + a0) / der
        polished.append(root)
    return polished
