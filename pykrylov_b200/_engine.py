"""Glue between the solver classes (reference API) and the device-resident iterations.

``resolve`` decides whether a (operator, preconditioner) pair can iterate entirely
on the GPU: the operator must be a :class:`CsrLinearOperator` and the
preconditioner absent or diagonal.  Anything else is a closure the device cannot
see; those go through the host-callback bridge (vectors stay in HBM, each
operator application crosses PCIe) -- never through a NumPy re-implementation of
the loop.
"""
import os

import numpy as np

from ._lib import KrylovDeviceError
from .device import DeviceSolver, DeviceVector, default_context, multi_axpy_dot, spmv_axpby_dot


class DevicePlan(object):
    def __init__(self, csr, precon_diag, precon_mode):
        self.csr = csr
        self.precon_diag = precon_diag
        self.precon_mode = precon_mode


def _diag_precon(precon, n):
    """(diag, mode) for preconditioners the device can apply itself, else None."""
    if precon is None:
        return None, 0
    from .linop import DiagonalOperator
    if isinstance(precon, DiagonalOperator):
        d = np.asarray(precon.diag)
        if d.dtype == np.float64 and d.shape == (n,):
            return d, 1                      # y = d .* r   (linop.py:473-503)
        return None
    d = getattr(precon, "diag", None)
    if d is not None and not hasattr(precon, "device_csr") and callable(precon):
        d = np.asarray(d)
        if d.dtype == np.float64 and d.shape == (n,) and _divides_by_diag(precon, d):
            return d, 2                      # y = r ./ d   (examples/bmark.py:14-22)
    return None


def _divides_by_diag(precon, d):
    """A callable with a ``.diag`` is only *assumed* to compute ``r / diag`` (bmark.py's
    DiagonalPrec); scaled Jacobi, a stored inverse diagonal or SSOR carry a ``.diag`` too.  So
    the device fast path is taken only after the callable has reproduced ``probe / diag`` to
    the bit on a random probe (IEEE division is correctly rounded: any honest ``r / diag``
    matches exactly); anything else stays an opaque callable and goes through the host bridge.
    The verdict is remembered per (object, diagonal buffer)."""
    key = (d.__array_interface__["data"][0], d.shape[0])
    seen = getattr(precon, "_kry_diag_probe", None)
    if seen is not None and seen[0] == key:
        return seen[1]
    probe = np.random.default_rng(20251017).standard_normal(d.shape[0])
    try:
        with np.errstate(all="ignore"):
            ok = bool(np.array_equal(np.asarray(precon(probe)), probe / d, equal_nan=True))
    except Exception:                        # noqa: BLE001 -- not callable that way: opaque
        ok = False
    try:
        precon._kry_diag_probe = (key, ok)
    except Exception:                        # noqa: BLE001 -- __slots__ etc.: probe again next time
        pass
    return ok


def resolve(op, precon, n):
    csr = getattr(op, "device_csr", None)
    if csr is None or csr.shape[0] != n or (not csr.sharded and csr.shape[1] != n):
        return None
    pd = _diag_precon(precon, n)
    if pd is None:
        return None
    return DevicePlan(csr, pd[0], pd[1])


def global_size(op, n):
    """Problem size the default iteration caps (2n, 5n) are derived from: the reference uses the
    system size; on a row shard that is the global row count, the same on every rank -- caps
    derived from the local slice would differ by rank and the ranks would latch `done` at
    different trips (and then wait for each other in the collectives forever)."""
    csr = getattr(op, "device_csr", None)
    if csr is not None and getattr(csr, "sharded", False):
        return int(csr.n_global)
    return n


def check_real(op, rhs):
    rtype = np.result_type(op.dtype, rhs.dtype)
    if np.issubdtype(rtype, np.complexfloating):
        raise TypeError("the device engine iterates in real fp64; complex systems are not supported")
    return np.result_type(rtype, np.float64) if not np.issubdtype(rtype, np.floating) else rtype


def make_solver(method, plan, context=None):
    """Device-resident iteration state for (operator, method).  The HBM slab of a solver
    (5-10 vectors) is kept on the operator and reused by later solves with the same method,
    so a solve does not pay a cudaMalloc/cudaFree of several hundred MB each time."""
    ctx = context or plan.csr.ctx
    cache = plan.csr.__dict__.setdefault("_solver_cache", {})
    S = cache.get(method)
    if S is None or not S._h.value or S.ctx is not ctx:
        S = DeviceSolver(ctx, method, plan.csr)
        cache[method] = S
    S.set_precon_diag(plan.precon_diag, plan.precon_mode)
    return S


def drive(S, check_interval, on_chunk=None, overlap=True):
    """Enqueue `check_interval` iterations at a time until the device latches done.
    One small D2H (status block + new history entries) per chunk.

    With ``overlap`` the host keeps one chunk in flight: chunk k+1 is enqueued before the
    status snapshot of chunk k is waited for and its history replayed, so the GPU never idles
    while Python works.  The chunk enqueued after the device latched `done` consists of no-op
    launches.  Callers whose ``on_chunk`` reads device vectors (store_iterates / store_resids)
    pass ``overlap=False``: they need the state of exactly the chunk they are told about."""
    st = S.status()
    if on_chunk is not None:
        on_chunk(st, S.drain_history(st))
    if st.done:
        return st
    if not overlap:
        while not st.done:
            S.iterate(check_interval)
            st = S.status()
            if on_chunk is not None:
                on_chunk(st, S.drain_history(st))
        return st
    slot = 0
    S.iterate(check_interval)
    S.status_enqueue(slot)
    while True:
        S.iterate(check_interval)                 # speculative: no-ops once `done` is latched
        S.status_enqueue(1 - slot)
        st = S.status_wait(slot)
        if on_chunk is not None:
            on_chunk(st, S.drain_history(st, nowait=True))
        if st.done:
            break
        slot = 1 - slot
    S.status_wait(1 - slot)                       # drain the speculative chunk before state is read
    return st


# ------------------------------------------------------------------ bridge
class HostBridge(object):
    """Vectors live in HBM; an opaque Python operator is applied by copying its
    argument to the host and its result back."""

    def __init__(self, n, context=None, op=None):
        csr = getattr(op, "device_csr", None)
        self.ctx = context or (csr.ctx if csr is not None else None) or default_context()
        self.n = n

    def vec(self, init=None):
        v = DeviceVector(self.ctx, self.n)
        if init is None:
            v.fill(0.0)
        else:
            v.upload(np.asarray(init, dtype=np.float64))
        return v

    def vec_n(self, n, init=None):
        v = DeviceVector(self.ctx, n)
        if init is None:
            v.fill(0.0)
        else:
            v.upload(np.asarray(init, dtype=np.float64))
        return v

    def apply(self, op, x_vec, out_vec, trans=False):
        """out = op * x (or op.T * x): a CUDA SpMV when the operator is a CSR in HBM,
        otherwise through the host closure.  Counts as one operator product."""
        csr = getattr(op, "device_csr", None)
        if csr is not None and not csr.sharded:
            if trans and not csr.symmetric:
                csr.build_transpose()
            csr.spmv(x_vec, out_vec, trans=trans)
            op._nMatvec += 1
            return out_vec
        chain = getattr(op, "device_apply", None)
        if chain is not None:                       # product of device operators: SpMVs back to back in HBM
            chain(x_vec, out_vec, trans=trans)
            op._nMatvec += 1
            return out_vec
        y = (op.T if trans else op) * x_vec.download()
        out_vec.upload(np.asarray(y, dtype=np.float64))
        return out_vec

    def apply_callable(self, fun, x_vec, out_vec):
        """out = fun(x) for an opaque host callable (lls preconditioners M, N)."""
        out_vec.upload(np.asarray(fun(x_vec.download()), dtype=np.float64))
        return out_vec

    def apply_precon(self, precon, x_vec, out_vec):
        out_vec.upload(np.asarray(precon * x_vec.download(), dtype=np.float64))
        return out_vec

    def fused(self, ops, dots=()):
        """kry_multi_axpy_dot + read back the dot results."""
        multi_axpy_dot(self.ctx, ops, dots, slot0=0)
        if dots:
            return self.ctx.scalars(0, len(dots))
        return ()


# ------------------------------------------------------------------ device-resident scalar planes
def plane_csr(op):
    """The unsharded device CSR behind `op` if the lls / SYMMLQ loops can run device-resident."""
    csr = getattr(op, "device_csr", None)
    if csr is None or getattr(csr, "sharded", False):
        return None
    return csr


class PlaneLoop(object):
    """Drives a solver whose vector work is stand-alone launches (SpMV, fused multi-AXPY + dot)
    and whose scalar recurrence lives in a device ScalarPlane: `trip()` enqueues one iteration
    without reading anything back; the status block and the history ring are read once per
    `check_interval` trips.  The plane's `done` flag gates every launch enqueued after the
    reference's stopping test fired, so the outcome does not depend on the interval."""

    def __init__(self, ctx, method):
        from .device import ScalarPlane
        cache = ctx.__dict__.setdefault("_scalar_planes", {})
        P = cache.get(method)
        if P is None or not P._h.value:
            P = ScalarPlane(ctx, method)
            cache[method] = P
        self.ctx, self.P = ctx, P
        self.use_graphs = True
        self.fuse = os.environ.get("KRY_LLS_FUSE", "1")
        self._tmp = {}

    def ops(self, ops, dots=(), step=None):
        """One fused vector pass; `step`: the phase of the recurrence that consumes its inner products,
        run in the launch's finalize instead of a kry_lls_step launch of its own."""
        if step is None or not dots or self.fuse == "0":
            multi_axpy_dot(self.ctx, ops, dots, slot0=0)
            if step is not None:
                self.P.step(step)
        else:
            multi_axpy_dot(self.ctx, ops, dots, plane=self.P, phase=step)

    # rows from which the recurrence phase after a fused SpMV runs as its own one-thread launch instead
    # of inside the SpMV's finalize: the in-kernel call costs the row kernel 8 registers (6 instead of 8
    # resident CTAs per SM; measured on B200), a separate launch costs ~3 us
    STEP_IN_KERNEL_BELOW = 1 << 18

    def spmv_ops(self, csr, x, op, step, dot_with=None, trans=False):
        """z = a*(A x) + b*w, the inner product dot_with . z (None: z . z) and phase `step`: one launch
        (two from STEP_IN_KERNEL_BELOW rows).  KRY_LLS_FUSE: 0 enqueues the three launches this replaces
        (the product into a temporary, the vector pass with its inner product, the step kernel), 2 / 3
        force the phase into / out of the SpMV launch (A/B measurements, tests)."""
        z = op["z"]
        if self.fuse == "0":
            t = self._tmp.get(z.n)
            if t is None:
                t = self._tmp[z.n] = DeviceVector(self.ctx, z.n)
            csr.spmv(x, t, trans=trans)
            multi_axpy_dot(self.ctx, [dict(op, u=t)], [(dot_with if dot_with is not None else z, z)], slot0=0)
            self.P.step(step)
        elif self.fuse == "2" or (self.fuse != "3" and z.n < self.STEP_IN_KERNEL_BELOW):
            spmv_axpby_dot(csr, x, op, dot_with=dot_with, trans=trans, plane=self.P, phase=step)
        else:
            spmv_axpby_dot(csr, x, op, dot=True, dot_with=dot_with, slot0=0, trans=trans)
            self.P.step(step)

    def run(self, trip, check_interval, on_chunk=None):
        """Returns (status, scalars) after the device latched `done`.  The gate is released on
        every exit path: later stand-alone launches of this context run unconditionally again."""
        from .device import LaunchGraph
        P = self.P
        interval = max(1, int(check_interval))
        graph, chunks = None, 0
        try:
            st, sc = P.status()
            while not st.done:
                if graph is not None:
                    graph.launch(interval)                 # one call per trip instead of ~10
                else:
                    for _ in range(interval):
                        trip()
                st, sc = P.status()
                if on_chunk is not None:
                    on_chunk(st, sc, P.drain_history(st))
                chunks += 1
                if (graph is None and chunks == 1 and not st.done and self.use_graphs and self.ctx.nranks == 1
                        and self.ctx.get_option(2)):
                    # the launch sequence of a trip is static and has now run un-captured: record it once
                    # (KRY_OPT_GRAPHS; not available on the host emulation)
                    try:
                        graph = LaunchGraph.capture(self.ctx, trip)
                    except KrylovDeviceError as exc:
                        # nothing of a failed capture has executed: keep enqueueing the same launches
                        import logging
                        logging.getLogger("pykrylov_b200").warning(
                            "trip graph capture failed (%s); the trips stay enqueued launch by launch", exc)
                        graph = None
                    if graph is None:
                        self.use_graphs = False
            return st, sc
        finally:
            P.release_gate()
            if graph is not None:
                graph._release()


def require_plan(method, op, precon, n):
    plan = resolve(op, precon, n)
    if plan is None:
        raise NotImplementedError(
            "%s currently iterates only on device operators (CsrLinearOperator / "
            "PysparseLinearOperator / CoordLinearOperator / linop_from_scipy) with no or a "
            "diagonal preconditioner; closure-defined operators are bridged for CG only" % method)
    return plan
