"""TEST-ONLY NumPy stand-in for pykrylov_b200._engine.HostBridge.

It restates the semantics of the C ABI calls the host-driven solvers (LSQR, SYMMLQ,
bridged CG) make -- kry_spmv and kry_multi_axpy_dot (ops applied in order per
element, un-fused multiply/add or divide, then the inner products) -- so that the
*host-side control logic* of those solvers can be tested on a machine without a GPU.
It is never importable from the product package."""
import numpy as np


class FakeVec(object):
    def __init__(self, n, init=None):
        self.a = np.zeros(n) if init is None else np.array(init, dtype=np.float64)
        self.n = n

    def download(self):
        return self.a.copy()

    def upload(self, v):
        self.a[:] = v
        return self

    def peek(self, i=0):
        return float(self.a[i])


class FakeBridge(object):
    def __init__(self, n, context=None, op=None):
        self.n = n

    def vec(self, init=None):
        return FakeVec(self.n, init)

    def vec_n(self, n, init=None):
        return FakeVec(n, init)

    def apply(self, op, x, out, trans=False):
        out.a[:] = (op.T if trans else op) * x.a.copy()
        return out

    def apply_precon(self, precon, x, out):
        out.a[:] = precon * x.a.copy()
        return out

    def apply_callable(self, fun, x, out):
        out.a[:] = fun(x.a.copy())
        return out

    @staticmethod
    def _term(c, v, neg, div):
        c = -c if neg else c
        return v / c if div else c * v

    def fused(self, ops, dots=()):
        sizes = {v.n for o in ops for v in (o["z"], o.get("u"), o.get("w")) if v is not None}
        sizes |= {v.n for pair in dots for v in pair}
        if len(sizes) > 1:               # kry_multi_axpy_dot: one launch = one vector length
            raise ValueError("kry_multi_axpy_dot: operand size mismatch %s" % sorted(sizes))
        for o in ops:
            u, w = o.get("u"), o.get("w")
            a, b = o.get("a", 1.0), o.get("b", 1.0)
            if u is not None and w is not None:
                r = self._term(a, u.a, o.get("a_neg"), o.get("a_div")) + self._term(b, w.a, o.get("b_neg"), o.get("b_div"))
            elif u is not None:
                r = self._term(a, u.a, o.get("a_neg"), o.get("a_div"))
            elif w is not None:
                r = self._term(b, w.a, o.get("b_neg"), o.get("b_div"))
            else:
                r = np.zeros_like(o["z"].a)
            o["z"].a[:] = r
        return [float(np.dot(u.a, w.a)) for u, w in dots]
