"""Linear operators: the plugin boundary the Krylov solvers iterate on.

API mirror of the reference's pykrylov/linop/linop.py (same constructors, same
``op * x`` / ``op.T`` / ``op.H`` protocol, same error types), re-implemented
around one idea: an operator is a triple of callables (matvec, transpose
matvec, adjoint matvec) and ``.T`` / ``.H`` are *views* that permute the triple
and point back to their origin (so ``A.T.T is A`` and, for real operators,
``A.H is A.T``, linop.py:148-170).

New in this engine: :class:`CsrLinearOperator`, an operator whose matrix lives
in HBM as CSR.  ``op * x`` on it runs the CUDA SpMV (host array in / out), and
the solvers recognise it and run their whole iteration device-resident.
"""
import logging

import numpy as np

from ..tools.types import allowed_types, complex_types, integer_types, real_types  # noqa: F401

__docformat__ = "restructuredtext"

null_log = logging.getLogger("linop")
null_log.setLevel(logging.INFO)
null_log.addHandler(logging.NullHandler())


class ShapeError(Exception):
    """Operand shapes do not fit the operator (reference linop.py:626-635)."""

    def __init__(self, value):
        Exception.__init__(self, value)
        self.value = value

    def __str__(self):
        return repr(self.value)


def _is_complex(dtype):
    try:
        return np.issubdtype(np.dtype(dtype), np.complexfloating)
    except TypeError:
        return dtype in complex_types


def _conj_wrap(fun):
    """x -> conj(fun(conj(x))): turns a transpose product into an adjoint one and back."""
    def wrapped(x):
        if np.iscomplexobj(x):
            return np.conjugate(fun(np.conjugate(x)))
        return np.conjugate(fun(x))
    return wrapped


class BaseLinearOperator(object):
    """Shape / type / counter bookkeeping shared by all operators (linop.py:14-104)."""

    def __init__(self, nargin, nargout, symmetric=False, hermitian=False, **kwargs):
        self._nargin = nargin
        self._nargout = nargout
        self._symmetric = symmetric
        self._hermitian = hermitian
        self._dtype = kwargs.get("dtype", np.float64)
        self._nMatvec = 0
        self.logger = kwargs.get("logger", null_log)
        self.logger.info("New linear operator with shape " + str(self.shape))

    nargin = property(lambda self: self._nargin, doc="The size of an input vector.")
    nargout = property(lambda self: self._nargout, doc="The size of an output vector.")
    symmetric = property(lambda self: self._symmetric, doc="Whether the operator is symmetric.")
    hermitian = property(lambda self: self._hermitian, doc="Whether the operator is Hermitian.")
    shape = property(lambda self: (self._nargout, self._nargin), doc="The shape of the operator.")
    nMatvec = property(lambda self: self._nMatvec,
                       doc="The number of products with vectors computed so far.")

    @property
    def dtype(self):
        "The data type of the operator."
        return self._dtype

    @dtype.setter
    def dtype(self, value):
        if value not in allowed_types:
            raise TypeError("Not a Numpy type")
        self._dtype = value

    def reset_counters(self):
        "Reset operator/vector product counter to zero."
        self._nMatvec = 0

    def __call__(self, *args, **kwargs):
        return self.__mul__(*args, **kwargs)

    def __mul__(self, x):
        raise NotImplementedError("Please subclass to implement __mul__.")

    def __repr__(self):
        kind = "Symmetric" if self.symmetric else "Unsymmetric"
        if self.hermitian:
            kind += " Hermitian"
        return "%s <%s> of type %s with shape (%d,%d)" % (
            kind, self.__class__.__name__, self.dtype, self.nargout, self.nargin)


class LinearOperator(BaseLinearOperator):
    """Operator defined by ``matvec`` and optionally ``matvec_transp`` / ``matvec_adj``
    (reference linop.py:107-452).  If ``symmetric`` the transpose is the operator
    itself; for real dtypes transpose and adjoint coincide; for complex dtypes a
    missing one is inferred from the other by conjugation."""

    def __init__(self, nargin, nargout, matvec, matvec_transp=None, matvec_adj=None, **kwargs):
        t_of = kwargs.pop("transpose_of", None)
        h_of = kwargs.pop("adjoint_of", None)
        kwargs.pop("conjugate_of", None)
        for name, link in (("transpose_of", t_of), ("adjoint_of", h_of)):
            if link is not None and not isinstance(link, BaseLinearOperator):
                raise ValueError("kwarg %s must be a BaseLinearOperator. Got %s"
                                 % (name, str(link.__class__)))
        super(LinearOperator, self).__init__(nargin, nargout, **kwargs)
        self._kw = dict((k, v) for k, v in kwargs.items() if k in ("dtype", "logger"))
        self._fun = matvec
        f_t = matvec if self.symmetric else matvec_transp
        f_h = matvec if self.hermitian else matvec_adj
        if not _is_complex(self.dtype):
            f_t = f_t if f_t is not None else f_h        # transpose == adjoint
            f_h = f_t
        else:
            if f_t is None and f_h is not None:
                f_t = _conj_wrap(f_h)
            elif f_h is None and f_t is not None:
                f_h = _conj_wrap(f_t)
        self._fun_t, self._fun_h = f_t, f_h
        self._T_view, self._H_view = t_of, h_of
        self._bar_view = None

    # ----------------------------------------------------------------- views
    def _view(self, fun, fun_t, fun_h, **links):
        return LinearOperator(self.nargout, self.nargin, fun, matvec_transp=fun_t, matvec_adj=fun_h,
                              **dict(self._kw, **links))

    @property
    def T(self):
        "The transpose operator."
        if self.symmetric:
            return self
        if self._T_view is None and self._fun_t is not None:
            cplx = _is_complex(self.dtype)
            if not cplx and self.hermitian:
                return self
            self._T_view = self._view(self._fun_t, self._fun,
                                      _conj_wrap(self._fun) if cplx else None, transpose_of=self)
            if not cplx:                      # real: A.H is A.T and (A.T).H is A
                self._H_view = self._T_view
                self._T_view._H_view = self
        return self._T_view

    @property
    def H(self):
        "The adjoint operator."
        if self.hermitian:
            return self
        if not _is_complex(self.dtype):
            return self.T
        if self._H_view is None and self._fun_h is not None:
            self._H_view = self._view(self._fun_h, _conj_wrap(self._fun), self._fun, adjoint_of=self)
        return self._H_view

    @property
    def bar(self):
        "The complex conjugate operator."
        return self.conjugate()

    def conjugate(self):
        "Return the complex conjugate operator (the operator itself when real)."
        if not _is_complex(self.dtype):
            return self
        if self._bar_view is None:
            self._bar_view = LinearOperator(self.nargin, self.nargout, _conj_wrap(self._fun),
                                            matvec_transp=self._fun_h, matvec_adj=self._fun_t,
                                            **self._kw)
            self._bar_view._bar_view = self
        return self._bar_view

    # -------------------------------------------------------------- products
    def _matvec(self, x):
        """Shape-checked call of the user closure (reference linop.py:271-298)."""
        x = np.asanyarray(x)
        nargout, nargin = self.shape
        try:
            x = x.reshape(nargin)
        except ValueError:
            raise ValueError("input array size incompatible with operator dimensions")
        y = self._fun(x)
        try:
            y = np.asanyarray(y).reshape(nargout)
        except ValueError:
            raise ValueError("output array size incompatible with operator dimensions")
        return y

    def rmatvec(self, x):
        "Product with the conjugate transpose (SciPy compatibility)."
        return self.H.__mul__(x)

    def to_array(self):
        "Dense matrix of the operator, one column per unit vector."
        nrow, ncol = self.shape
        out = np.empty((nrow, ncol), dtype=self.dtype)
        e = np.zeros(ncol, dtype=self.dtype)
        for j in range(ncol):
            e[j] = 1
            out[:, j] = self * e
            e[j] = 0
        return out

    full = to_array

    def _times_scalar(self, alpha):
        rtype = np.result_type(self.dtype, type(alpha))
        if alpha == 0:
            return ZeroOperator(self.nargin, self.nargout, dtype=rtype)
        return LinearOperator(self.nargin, self.nargout,
                              matvec=lambda v: alpha * self(v),
                              matvec_transp=lambda v: alpha * self.T(v),
                              matvec_adj=lambda v: np.conjugate(alpha) * self.H(v),
                              symmetric=self.symmetric,
                              hermitian=(not _is_complex(rtype)) and self.hermitian,
                              dtype=rtype)

    def _times_operator(self, other):
        if self.nargin != other.nargout:
            raise ShapeError("Cannot multiply operators together")
        return LinearOperator(other.nargin, self.nargout,
                              matvec=lambda v: self(other(v)),
                              matvec_transp=lambda v: other.T(self.T(v)),
                              matvec_adj=lambda v: other.H(self.H(v)),
                              symmetric=False, hermitian=False,
                              dtype=np.result_type(self.dtype, other.dtype))

    def _times_vector(self, x):
        self._nMatvec += 1
        return self._matvec(x).astype(np.result_type(self.dtype, x.dtype))

    def __mul__(self, x):
        if np.isscalar(x):
            return self._times_scalar(x)
        if isinstance(x, BaseLinearOperator):
            return self._times_operator(x)
        if isinstance(x, np.ndarray):
            return self._times_vector(x)
        raise ValueError("Cannot multiply")

    def __rmul__(self, x):
        if np.isscalar(x):
            return self._times_scalar(x)
        if isinstance(x, BaseLinearOperator):
            return x._times_operator(self)
        raise ValueError("Cannot multiply")

    def _combine(self, other, sign):
        if not isinstance(other, BaseLinearOperator):
            raise ValueError("Cannot add")
        if self.shape != other.shape:
            raise ShapeError("Cannot add")
        if isinstance(other, CsrLinearOperator) and not isinstance(self, CsrLinearOperator):
            dev = other._device_combine_reversed(self, sign)        # D +- A with A in HBM
            if dev is not None:
                return dev
        return LinearOperator(self.nargin, self.nargout,
                              matvec=lambda v: self(v) + sign * other(v),
                              matvec_transp=lambda v: self.T(v) + sign * other.T(v),
                              matvec_adj=lambda v: self.H(v) + sign * other.H(v),
                              symmetric=self.symmetric and other.symmetric,
                              hermitian=self.hermitian and other.hermitian,
                              dtype=np.result_type(self.dtype, other.dtype))

    def __add__(self, other):
        return self._combine(other, 1)

    def __sub__(self, other):
        return self._combine(other, -1)

    def __neg__(self):
        return self * (-1)

    def __truediv__(self, other):
        if not np.isscalar(other):
            raise ValueError("Cannot divide")
        return self * (1. / other)

    __div__ = __truediv__

    def __pow__(self, k):
        if not isinstance(k, int):
            raise ValueError("Can only raise to integer power")
        if k < 0:
            raise ValueError("Can only raise to nonnegative power")
        if self.nargin != self.nargout:
            raise ShapeError("Can only raise square operators to a power")
        if k == 0:
            return IdentityOperator(self.nargin)
        return self if k == 1 else self * self ** (k - 1)


def _drop(kwargs, *names):
    for name in names:
        kwargs.pop(name, None)
    return kwargs


class IdentityOperator(LinearOperator):
    """The identity of order ``nargin`` (reference linop.py:455-470)."""

    def __init__(self, nargin, **kwargs):
        _drop(kwargs, "symmetric", "hermitian", "matvec")
        super(IdentityOperator, self).__init__(nargin, nargin, matvec=lambda x: x,
                                               symmetric=True, hermitian=True, **kwargs)

    def _times_scalar(self, alpha):
        # sigma * I as a diagonal operator: the same values as the reference's closure
        # (sigma * x either way) in a form the device operator algebra can take in (A + sigma*I)
        if alpha != 0 and isinstance(alpha, (int, float, np.integer, np.floating)) \
                and not _is_complex(self.dtype):
            return DiagonalOperator(np.full(self.nargin, float(alpha)))
        return super(IdentityOperator, self)._times_scalar(alpha)


class DiagonalOperator(LinearOperator):
    """diag(d); the dtype is that of ``d`` (reference linop.py:473-516)."""

    def __init__(self, diag, **kwargs):
        _drop(kwargs, "symmetric", "hermitian", "matvec", "matvec_adj", "dtype")
        diag = np.asarray(diag)
        if diag.ndim != 1:
            raise ValueError("Input must be 1-d array")
        self._diag = diag.copy()
        super(DiagonalOperator, self).__init__(diag.shape[0], diag.shape[0],
                                               matvec=lambda x: diag * x,
                                               matvec_adj=lambda x: diag.conjugate() * x,
                                               symmetric=True,
                                               hermitian=not _is_complex(diag.dtype),
                                               dtype=diag.dtype, **kwargs)

    @property
    def diag(self):
        "A reference to the diagonal of the operator."
        return self._diag

    def __abs__(self):
        return DiagonalOperator(np.abs(self._diag))

    def _sqrt(self):
        if not _is_complex(self.dtype) and np.any(self._diag < 0):
            raise ValueError("Math domain error")
        return DiagonalOperator(np.sqrt(self._diag))


class ZeroOperator(LinearOperator):
    """The ``nargout`` x ``nargin`` zero map (reference linop.py:519-557)."""

    def __init__(self, nargin, nargout, **kwargs):
        _drop(kwargs, "matvec", "matvec_transp")

        def zeros_for(n_in, n_out):
            def fun(x):
                if x.shape != (n_in,):
                    raise ValueError("Input has shape %s instead of (%d,)" % (str(x.shape), n_in))
                return np.zeros(n_out, dtype=np.result_type(self.dtype, x.dtype))
            return fun

        back = zeros_for(nargout, nargin)
        super(ZeroOperator, self).__init__(nargin, nargout, matvec=zeros_for(nargin, nargout),
                                           matvec_transp=back, matvec_adj=back,
                                           symmetric=(nargin == nargout), **kwargs)

    def __abs__(self):
        return self

    def _sqrt(self):
        return self


def _restricted(op, rows, cols, symmetric, hermitian):
    """Operator x -> (op * scatter(x, cols))[rows] and its transposes."""
    nrow_full, ncol_full = op.shape

    def make(which, n_embed, idx_in, idx_out):
        def fun(x):
            z = np.zeros(n_embed, dtype=x.dtype)
            z[idx_in] = x[:]
            return (which() * z)[idx_out]
        return fun

    return LinearOperator(len(cols), len(rows),
                          matvec=make(lambda: op, ncol_full, cols, rows),
                          matvec_transp=make(lambda: op.T, nrow_full, rows, cols),
                          matvec_adj=make(lambda: op.H, nrow_full, rows, cols),
                          symmetric=symmetric, hermitian=hermitian, dtype=op.dtype)


def ReducedLinearOperator(op, row_indices, col_indices):
    """Restrict inputs to ``col_indices`` and outputs to ``row_indices`` (linop.py:560-590)."""
    return _restricted(op, row_indices, col_indices, False, False)


def SymmetricallyReducedLinearOperator(op, indices):
    """Restrict inputs and outputs to the same ``indices`` (linop.py:593-623)."""
    return _restricted(op, indices, indices, op.symmetric, op.hermitian)


# ------------------------------------------------------------------ device
class CsrLinearOperator(LinearOperator):
    """Operator whose matrix is a CSR resident in HBM (pykrylov_b200.device.DeviceCsr).

    ``op * x`` uploads x, runs the CUDA SpMV and downloads the result -- the drop-in
    for the reference's closure call chain (linop.py:362-369 -> :356-360 -> :271-298
    -> user ``matvec``).  The Krylov solvers detect the ``device_csr`` attribute and
    keep the entire iteration on the GPU instead of calling ``op * x`` per step.
    """

    def __init__(self, device_csr, **kwargs):
        self.device_csr = device_csr
        nrow, ncol = device_csr.shape
        if getattr(device_csr, "sharded", False):
            ncol = nrow                       # SPMD view: local slice in, local slice out
        _drop(kwargs, "symmetric", "matvec", "matvec_transp", "dtype")

        def matvec(x):
            if x.shape != (ncol,):
                raise ShapeError("Input has shape %s instead of (%d,)" % (str(x.shape), ncol))
            return device_csr.matvec(x)

        def matvec_transp(y):
            if y.shape != (nrow,):
                raise ShapeError("Input has shape %s instead of (%d,)" % (str(y.shape), nrow))
            if not device_csr.symmetric:
                device_csr.build_transpose()          # device-side, once
            return device_csr.matvec(y, trans=True)

        super(CsrLinearOperator, self).__init__(ncol, nrow, matvec=matvec,
                                                matvec_transp=matvec_transp,
                                                symmetric=device_csr.symmetric,
                                                hermitian=device_csr.symmetric,
                                                dtype=np.float64, **kwargs)

    @property
    def T(self):
        t = LinearOperator.T.fget(self)
        if t is not self and t is not None and not hasattr(t, "device_csr_transpose_of"):
            t.device_csr_transpose_of = self.device_csr
        return t

    def diagonal(self):
        return self.device_csr.diagonal()

    # -- operator algebra that stays in HBM (reference linop.py:307-345, 378-410 builds host
    #    closures; a closure handed to a solver would leave the device loop and cross PCIe at every
    #    product).  The combined operator is a new CSR: per row the (scaled) entries of the first
    #    operand, then those of the second, then a diagonal term.  A +- sigma*I, A +- D and -A equal
    #    the reference's expression bit for bit (the appended entry is added last, like the second
    #    summand; negation is exact); alpha*A and A + B differ from it by rounding only (one row
    #    sum instead of alpha*(sum) / two sums).
    def _algebra_ok(self):
        return not getattr(self.device_csr, "sharded", False)

    @staticmethod
    def _diag_of(op):
        """Real fp64 diagonal of an Identity / Diagonal operator, else None."""
        if isinstance(op, IdentityOperator) and not _is_complex(op.dtype):
            return np.ones(op.nargin)
        if isinstance(op, DiagonalOperator):
            d = np.asarray(op.diag)
            if d.dtype == np.float64:
                return d
        return None

    def _times_scalar(self, alpha):
        if alpha == 0 or not self._algebra_ok() or not isinstance(alpha, (int, float, np.integer, np.floating)):
            return super(CsrLinearOperator, self)._times_scalar(alpha)
        return CsrLinearOperator(self.device_csr.combine(alpha=float(alpha)))

    def _combine(self, other, sign):
        if not isinstance(other, BaseLinearOperator):
            raise ValueError("Cannot add")
        if self.shape != other.shape:
            raise ShapeError("Cannot add")
        if self._algebra_ok():
            if isinstance(other, CsrLinearOperator) and other._algebra_ok() \
                    and other.device_csr.ctx is self.device_csr.ctx:
                return CsrLinearOperator(self.device_csr.combine(1.0, other.device_csr, float(sign)))
            d = self._diag_of(other)
            if d is not None:
                return CsrLinearOperator(self.device_csr.combine(1.0, diag=d, gamma=float(sign), symmetric=self.symmetric))
        return super(CsrLinearOperator, self)._combine(other, sign)

    def _device_combine_reversed(self, left, sign):
        """left +- self for an Identity / Diagonal `left`: (+-1)*self + diag(d)."""
        d = self._diag_of(left)
        if d is None or not self._algebra_ok():
            return None
        return CsrLinearOperator(self.device_csr.combine(float(sign), diag=d, gamma=1.0, symmetric=self.symmetric))

    def _times_operator(self, other):
        if self.nargin != other.nargout:
            raise ShapeError("Cannot multiply operators together")
        if isinstance(other, (CsrLinearOperator, DeviceChainOperator)) and self._algebra_ok():
            return DeviceChainOperator(DeviceChainOperator.links(self) + DeviceChainOperator.links(other))
        return super(CsrLinearOperator, self)._times_operator(other)

    def to_array(self):
        "Dense matrix of the operator, scattered on the device (one D2H instead of ncol products)."
        if not self._algebra_ok():
            return super(CsrLinearOperator, self).to_array()
        return self.device_csr.to_dense()

    full = to_array


class DeviceChainOperator(LinearOperator):
    """Product A_1 A_2 ... A_k of operators that live in HBM: ``op * x`` uploads x once, applies
    the CUDA SpMVs back to back on device vectors and downloads the result once (the reference
    composes host closures, linop.py:320-331).  Solvers run it through the host-callback bridge
    without leaving the device (``device_apply``)."""

    @staticmethod
    def links(op):
        return list(op._links) if isinstance(op, DeviceChainOperator) else [op.device_csr]

    def __init__(self, links):
        self._links = list(links)
        ctx = links[0].ctx
        if any(m.ctx is not ctx for m in links):
            raise ValueError("operators of a product must live on the same device context")
        nargout, nargin = links[0].shape[0], links[-1].shape[1]

        def run(x, trans):
            from ..device import DeviceVector
            xv = DeviceVector(ctx, x.shape[0]).upload(x)
            out = self.device_apply(xv, None, trans=trans)
            return out.download()

        def matvec(x):
            if x.shape != (nargin,):
                raise ShapeError("Input has shape %s instead of (%d,)" % (str(x.shape), nargin))
            return run(x, False)

        def matvec_transp(y):
            if y.shape != (nargout,):
                raise ShapeError("Input has shape %s instead of (%d,)" % (str(y.shape), nargout))
            return run(y, True)

        super(DeviceChainOperator, self).__init__(nargin, nargout, matvec=matvec, matvec_transp=matvec_transp,
                                                  symmetric=False, hermitian=False, dtype=np.float64)

    def device_apply(self, x_vec, out_vec=None, trans=False):
        """out = (A_1 ... A_k) x, or its transpose applied to x, on device vectors."""
        from ..device import DeviceVector
        order = self._links if trans else self._links[::-1]
        cur = x_vec
        for i, m in enumerate(order):
            if trans and not m.symmetric:
                m.build_transpose()
            n_out = m.shape[1] if trans else m.shape[0]
            dst = out_vec if (i == len(order) - 1 and out_vec is not None) else DeviceVector(m.ctx, n_out)
            m.spmv(cur, dst, trans=trans)
            cur = dst
        return cur

    def _times_operator(self, other):
        if self.nargin != other.nargout:
            raise ShapeError("Cannot multiply operators together")
        if isinstance(other, (CsrLinearOperator, DeviceChainOperator)) and \
                not (isinstance(other, CsrLinearOperator) and not other._algebra_ok()):
            return DeviceChainOperator(self._links + DeviceChainOperator.links(other))
        return super(DeviceChainOperator, self)._times_operator(other)


def csr_operator(shape, indptr, indices, data, symmetric=False, context=None):
    """Device operator from host CSR arrays (int32/int64 indices, fp64 values)."""
    from ..device import DeviceCsr, default_context
    ctx = context or default_context()
    return CsrLinearOperator(DeviceCsr.from_arrays(ctx, shape, indptr, indices, data,
                                                   symmetric=symmetric))


def linop_from_scipy(M, symmetric=False, context=None):
    """Device operator from a scipy.sparse matrix (converted to sorted CSR)."""
    M = M.tocsr()
    M.sort_indices()
    return csr_operator(M.shape, M.indptr, M.indices, M.data, symmetric=symmetric, context=context)


def _coo_to_csr_in_arrival_order(vals, rows, cols, nrows, symmetric):
    """Expand (optionally one-triangle) COO triplets to CSR keeping, inside every
    row, the order in which the reference's loop (linop.py:657-664) would have
    accumulated them -- so the device row sums are bit-identical to that loop."""
    rows = np.asarray(rows).astype(np.int64)
    cols = np.asarray(cols).astype(np.int64)
    vals = np.asarray(vals)
    if symmetric:
        off = rows != cols
        seq = np.arange(len(vals)) * 2
        r = np.concatenate([rows, cols[off]])
        c = np.concatenate([cols, rows[off]])
        v = np.concatenate([vals, vals[off]])
        order_key = np.concatenate([seq, seq[off] + 1])
    else:
        r, c, v = rows, cols, vals
        order_key = np.arange(len(vals))
    perm = np.lexsort((order_key, r))
    r, c, v = r[perm], c[perm], v[perm]
    indptr = np.zeros(nrows + 1, dtype=np.int64)
    np.add.at(indptr, r + 1, 1)
    return np.cumsum(indptr), c, v


def CoordLinearOperator(vals, rows, cols, nargin=0, nargout=0, symmetric=False, context=None):
    """Operator from coordinate triplets (reference linop.py:638-685).  With
    ``symmetric=True`` the triplets describe one triangle.  Real fp64 data is
    assembled into a device CSR; other dtypes keep a host closure."""
    if nargin == 0:
        nargin = cols.max()           # (sic) reference default, linop.py:646
    if nargout == 0:
        nargout = rows.max()          # (sic) linop.py:647
    nargin, nargout = int(nargin), int(nargout)
    rows_i, cols_i = np.asarray(rows).astype(np.int64), np.asarray(cols).astype(np.int64)
    vals = np.asarray(vals)
    if vals.dtype == np.float64:
        # assembled in HBM (csrc/assemble.cu): stable sort of the triplets by row, symmetric
        # expansion and the range check all run on the device; A^T comes from the same triplets
        # in the order of the reference's matvec_transp
        from ..device import DeviceCsr, default_context
        from .._lib import KrylovDeviceError, KRY_ERR_INVALID
        ctx = context or default_context()
        try:
            csr = DeviceCsr.from_coo(ctx, (nargout, nargin), rows_i, cols_i, vals, symmetric=symmetric,
                                     build_transpose=not symmetric)
        except KrylovDeviceError as exc:
            if exc.status == KRY_ERR_INVALID:
                raise IndexError("coordinate index out of bounds for a (%d,%d) operator" % (nargout, nargin))
            raise
        return CsrLinearOperator(csr)
    if len(vals) and (rows_i.min() < 0 or cols_i.min() < 0 or rows_i.max() >= nargout
                      or cols_i.max() >= nargin
                      or (symmetric and (rows_i.max() >= nargin or cols_i.max() >= nargout))):
        raise IndexError("coordinate index out of bounds for a (%d,%d) operator" % (nargout, nargin))

    def scatter(n_in, n_out, out_idx, in_idx):
        def fun(x):
            if x.shape != (n_in,):
                raise ShapeError("Input has shape %s instead of (%d,)" % (str(x.shape), n_in))
            y = np.zeros(n_out, dtype=np.result_type(x.dtype, vals.dtype))
            np.add.at(y, out_idx, vals * x[in_idx])
            if symmetric:
                off = out_idx != in_idx
                np.add.at(y, in_idx[off], vals[off] * x[out_idx[off]])
            return y
        return fun

    return LinearOperator(nargin, nargout, matvec=scatter(nargin, nargout, rows_i, cols_i),
                          matvec_transp=scatter(nargout, nargin, cols_i, rows_i),
                          symmetric=symmetric, dtype=vals.dtype)


def PysparseLinearOperator(A, context=None):
    """Operator from a Pysparse-style matrix (reference linop.py:688-720).

    Matrices that can export CSR (the bundled ``pysparse`` shim, anything with
    ``to_csr_arrays()``) become device operators; any other object is wrapped as a
    host closure exactly like the reference does."""
    nargout, nargin = A.shape
    try:
        symmetric = A.issym
    except AttributeError:
        symmetric = A.isSymmetric()
    if hasattr(A, "to_csr_arrays"):
        indptr, indices, data = A.to_csr_arrays()
        return csr_operator((nargout, nargin), indptr, indices, data, symmetric=bool(symmetric),
                            context=context)

    def matvec(x):
        if x.shape != (nargin,):
            raise ShapeError("Input has shape %s instead of (%d,)" % (str(x.shape), nargin))
        if hasattr(A, "__mul__"):
            return A * x
        out = np.empty(nargout)
        A.matvec(x, out)
        return out

    def matvec_transp(y):
        if y.shape != (nargout,):
            raise ShapeError("Input has shape %s instead of (%d,)" % (str(y.shape), nargout))
        if hasattr(A, "__rmul__"):
            return y * A
        out = np.empty(nargin)
        A.matvec_transp(y, out)
        return out

    return LinearOperator(nargin, nargout, matvec=matvec, matvec_transp=matvec_transp,
                          symmetric=symmetric)


def linop_from_ndarray(A, symmetric=False, **kwargs):
    """Operator from a dense ndarray (reference linop.py:723-745)."""
    hermitian = kwargs.get("hermitian", symmetric)
    if _is_complex(A.dtype):
        return LinearOperator(A.shape[1], A.shape[0], lambda v: np.dot(A, v),
                              matvec_transp=lambda u: np.dot(A.T, u),
                              matvec_adj=lambda w: np.dot(A.conjugate().T, w),
                              symmetric=symmetric, hermitian=hermitian, dtype=A.dtype)
    if symmetric ^ hermitian:
        raise ValueError("For non-complex operators, transpose = adjoint.")
    both = symmetric or hermitian
    return LinearOperator(A.shape[1], A.shape[0], lambda v: np.dot(A, v),
                          matvec_transp=lambda u: np.dot(A.T, u),
                          symmetric=both, hermitian=both, dtype=A.dtype)


def sqrt(op):
    """Operator square root where defined (not element-wise; reference linop.py:748-754)."""
    return op._sqrt()
