"""Thin object layer over the C ABI: Context, DeviceVector, DeviceCsr, DeviceSolver.

Host code stays pure Python + NumPy; every array crossing this layer is copied
into / out of HBM by libkrylov_b200 (the library never keeps a host pointer).
"""
import ctypes as C
import os
import weakref

import numpy as np

from . import _lib as L
from ._lib import call

__all__ = ["Context", "DeviceVector", "DeviceCsr", "DeviceSolver", "default_context", "device_count"]


def device_count():
    n = C.c_int(0)
    call("kry_device_count", C.byref(n))
    return n.value


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


# ------------------------------------------------------------------ pinned results
class _PinnedBlock(object):
    """One page-locked host block.  NumPy arrays made from it (``np.asarray(block)``)
    keep it alive through ``.base``; when the last of them is collected the block goes
    back to its pool instead of being unpinned."""
    __slots__ = ("ptr", "nbytes", "pool", "__array_interface__", "__weakref__")

    def __init__(self, ptr, nbytes, count, pool):
        self.ptr, self.nbytes, self.pool = ptr, nbytes, pool
        self.__array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 3}

    def __del__(self):
        try:
            self.pool._give_back(self.ptr, self.nbytes)
        except Exception:
            pass


class PinnedPool(object):
    """Recycles page-locked result buffers.

    A device->host copy into pageable memory runs at a fraction of the PCIe rate (the
    driver bounces it through its own staging buffer, and a fresh ``np.empty`` page-faults
    on first touch): for the 80 MB solution of a 10^7-row solve that is ~10 ms against
    ~1.5 ms into pinned memory.  Pinning itself is slow, so blocks are pooled: the arrays a
    solve returns (``bestSolution``, ``op * x``) are views of pinned blocks that return to
    the pool when the caller drops them.  Small results, and anything beyond the caps, use
    ordinary ``np.empty`` arrays.
    """

    def __init__(self, alloc=None, free=None, min_bytes=1 << 20, live_cap=8 << 30, idle_cap=2 << 30):
        self._alloc = alloc or self._cuda_alloc
        self._free = free or self._cuda_free
        self.min_bytes, self.live_cap, self.idle_cap = int(min_bytes), int(live_cap), int(idle_cap)
        self._idle = {}              # nbytes -> [ptr, ...]
        self.live_bytes = 0          # handed out
        self.idle_bytes = 0          # parked in the pool
        self.hits = self.misses = 0

    @staticmethod
    def _cuda_alloc(nbytes):
        p = C.c_void_p()
        if L.lib.kry_host_alloc(int(nbytes), C.byref(p)) != L.KRY_OK or not p.value:
            return None
        return p.value

    @staticmethod
    def _cuda_free(ptr):
        L.lib.kry_host_free(C.c_void_p(ptr))

    def empty(self, n):
        """Uninitialised fp64[n]; pinned when it pays and the caps allow, else pageable."""
        n = int(n)
        nbytes = 8 * n
        if nbytes < self.min_bytes or self.live_bytes + nbytes > self.live_cap:
            return np.empty(n, dtype=np.float64)
        size = (nbytes + 0xFFFFF) & ~0xFFFFF            # 1 MiB classes
        stack = self._idle.get(size)
        if stack:
            ptr = stack.pop()
            self.idle_bytes -= size
            self.hits += 1
        else:
            ptr = self._alloc(size)
            if ptr is None:
                return np.empty(n, dtype=np.float64)
            self.misses += 1
        self.live_bytes += size
        return np.asarray(_PinnedBlock(ptr, size, n, self))

    def _give_back(self, ptr, size):
        self.live_bytes -= size
        if self.idle_bytes + size > self.idle_cap:
            self._free(ptr)
            return
        self._idle.setdefault(size, []).append(ptr)
        self.idle_bytes += size

    def trim(self):
        """Unpin everything that is parked."""
        for stack in self._idle.values():
            while stack:
                self._free(stack.pop())
        self.idle_bytes = 0


result_pool = PinnedPool()


class Context(object):
    """One CUDA device + stream + reduction workspace (``kry_ctx``)."""

    def __init__(self, device=None):
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
            n = device_count()
            if n > 0:
                device %= n
        self._h = L.handle()
        call("kry_ctx_create", int(device), C.byref(self._h))
        self.device = int(device)
        self.nranks, self.rank = 1, 0
        self._children = weakref.WeakSet()     # vectors / operators / solvers living on this context

    # -- lifetime
    def _adopt(self, obj):
        self._children.add(obj)

    def close(self):
        """Destroy the context after every object that still lives on it."""
        if getattr(self, "_h", None) is not None and self._h.value:
            order = {"LaunchGraph": -1, "DeviceSolver": 0, "ScalarPlane": 0, "DeviceCsr": 1, "DeviceVector": 2}
            for obj in sorted(list(self._children), key=lambda o: order.get(type(o).__name__, 3)):
                obj._release()
            L.lib.kry_ctx_destroy(self._h)
            self._h = L.handle()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- misc
    def sync(self):
        call("kry_ctx_sync", self._h)

    def props(self):
        p = (C.c_int64 * 6)()
        call("kry_ctx_props", self._h, p)
        keys = ("sm_count", "total_bytes", "free_bytes", "cc", "l2_bytes", "smem_optin")
        return dict(zip(keys, [int(v) for v in p]))

    def timer_start(self):
        call("kry_timer_start", self._h)

    def timer_stop(self):
        ms = C.c_double(0.0)
        call("kry_timer_stop", self._h, C.byref(ms))
        return ms.value

    def flush_l2(self):
        call("kry_flush_l2", self._h)

    def launch_count(self):
        n = C.c_int64(0)
        call("kry_launch_count", self._h, C.byref(n))
        return n.value

    def halo_trace(self):
        """KRY_HALO_TRACE diagnostics (kry_halo_trace_read); zeros unless tracing is on."""
        out = (C.c_uint64 * 16)()
        call("kry_halo_trace_read", self._h, out)
        return [int(v) for v in out]

    def set_option(self, option, value):
        call("kry_ctx_set_option", self._h, int(option), int(value))

    def get_option(self, option):
        v = C.c_int(0)
        call("kry_ctx_get_option", self._h, int(option), C.byref(v))
        return v.value

    def prof_enable(self, max_samples):
        call("kry_prof_enable", self._h, int(max_samples))

    def prof_read(self):
        """(samples, total_ms) of the fused SpMV+dot launches since the last read."""
        n, ms = C.c_int64(0), C.c_double(0.0)
        call("kry_prof_read", self._h, C.byref(n), C.byref(ms))
        return n.value, ms.value

    def pinned_array(self, n):
        """fp64 NumPy array backed by page-locked host memory (DMA-able staging); the
        memory is unpinned/pooled when the array and all its views are collected."""
        pool = result_pool
        n = int(n)
        size = (8 * max(n, 1) + 0xFFFFF) & ~0xFFFFF
        ptr = pool._alloc(size)
        if ptr is None:
            raise L.KrylovDeviceError(L.KRY_ERR_NOMEM, "pinned host allocation of %d bytes failed: %s"
                                      % (size, L.last_error()))
        pool.live_bytes += size
        return np.asarray(_PinnedBlock(ptr, size, n, pool))

    def scalars(self, first=0, count=L.KRY_NUM_SLOTS):
        out = (C.c_double * count)()
        call("kry_scalars_read", self._h, first, count, out)
        return np.array(out[:count])

    def set_scalars(self, first, values):
        values = [float(v) for v in values]
        buf = (C.c_double * len(values))(*values)
        call("kry_scalars_write", self._h, first, len(values), buf)

    # -- factories
    def vector(self, n_or_array):
        if np.isscalar(n_or_array):
            return DeviceVector(self, int(n_or_array))
        a = _f64(n_or_array)
        v = DeviceVector(self, a.shape[0])
        v.upload(a)
        return v

    # -- multi-GPU
    def comm_init(self, nranks, rank, unique_id):
        call("kry_comm_init", self._h, int(nranks), int(rank), C.c_char_p(unique_id))
        self.nranks, self.rank = int(nranks), int(rank)

    def barrier(self):
        call("kry_comm_barrier", self._h)

    def allreduce(self, values, op="sum"):
        vals = [float(v) for v in np.atleast_1d(values)]
        buf = (C.c_double * len(vals))(*vals)
        call("kry_comm_allreduce_host", self._h, buf, len(vals), 1 if op == "max" else 0)
        return np.array(buf[:len(vals)])

    def allgather_bytes(self, payload):
        payload = bytes(payload)
        out = C.create_string_buffer(len(payload) * self.nranks)
        call("kry_comm_allgather_host", self._h, C.c_char_p(payload), out, len(payload))
        raw = out.raw
        return [raw[i * len(payload):(i + 1) * len(payload)] for i in range(self.nranks)]


_default = None


def default_context():
    """Process-wide context on cuda:LOCAL_RANK (created on first use)."""
    global _default
    if _default is None:
        _default = Context()
    return _default


class DeviceVector(object):
    """fp64 vector resident in HBM (``kry_vec``)."""

    def __init__(self, ctx, n, capacity=None):
        self.ctx = ctx
        self.n = int(n)
        self.capacity = self.n if capacity is None else int(capacity)
        self._h = L.handle()
        call("kry_vec_create_cap", ctx._h, self.n, self.capacity, C.byref(self._h))
        ctx._adopt(self)

    def _release(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            L.lib.kry_vec_destroy(self._h)
            self._h = L.handle()

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def __len__(self):
        return self.n

    def upload(self, a):
        a = _f64(a)
        call("kry_vec_upload", self._h, _ptr(a), a.shape[0])
        return self

    def download(self, out=None):
        if out is None:
            out = result_pool.empty(self.n)
        call("kry_vec_download", self._h, _ptr(out), out.shape[0])
        return out

    def fill(self, value):
        call("kry_vec_fill", self._h, float(value))
        return self

    def peek(self, index=0):
        """One entry (for the log lines that print x[0])."""
        out = C.c_double(0.0)
        call("kry_vec_read", self._h, int(index), 1, C.byref(out))
        return out.value

    def copy_from(self, other):
        call("kry_vec_copy", self._h, other._h)
        return self


class DeviceCsr(object):
    """CSR operator resident in HBM (``kry_csr``): int32 indices, fp64 values."""

    def __init__(self, ctx, handle, symmetric):
        self.ctx = ctx
        self._h = handle
        self.symmetric = bool(symmetric)
        nr, nc, nz = C.c_int64(), C.c_int64(), C.c_int64()
        call("kry_csr_shape", self._h, C.byref(nr), C.byref(nc), C.byref(nz))
        self.shape = (nr.value, nc.value)
        self.nnz = nz.value
        self.sharded = False          # set by shard_finalize(): columns address [local | halo]
        self.n_global = nr.value
        ctx._adopt(self)

    def _release(self):
        for S in list(self.__dict__.get("_solver_cache", {}).values()):
            S._release()                      # solvers reference the operator: free them first
        self.__dict__.pop("_solver_cache", None)
        if getattr(self, "_h", None) is not None and self._h.value:
            L.lib.kry_csr_destroy(self._h)
            self._h = L.handle()

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    # -- constructors
    @classmethod
    def from_arrays(cls, ctx, shape, indptr, indices, data, symmetric=False, build_transpose=False):
        indptr = np.ascontiguousarray(indptr)
        if indptr.dtype != np.int32:
            if len(indptr) and int(indptr[-1]) >= 2 ** 31 - 2 ** 20:
                raise L.KrylovDeviceError(L.KRY_ERR_UNSUPPORTED, "nnz exceeds the int32 index space")
            indptr = indptr.astype(np.int32)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        data = _f64(data)
        flags = (L.KRY_CSR_SYMMETRIC if symmetric else 0) | (L.KRY_CSR_BUILD_TRANSPOSE if build_transpose else 0)
        h = L.handle()
        call("kry_csr_create", ctx._h, int(shape[0]), int(shape[1]), int(data.shape[0]),
             _ptr(indptr), _ptr(indices), _ptr(data), flags, C.byref(h))
        return cls(ctx, h, symmetric)

    @classmethod
    def from_coo(cls, ctx, shape, rows, cols, vals, symmetric=False, build_transpose=False):
        """Assembled on the device from coordinate triplets (``kry_csr_create_coo``): inside a row
        the entries keep their arrival order; ``symmetric`` expands one stored triangle."""
        rows = np.ascontiguousarray(rows, dtype=np.int32)
        cols = np.ascontiguousarray(cols, dtype=np.int32)
        vals = _f64(vals)
        if not (rows.shape == cols.shape == vals.shape and rows.ndim == 1):
            raise ValueError("rows, cols and vals must be 1-d arrays of the same length")
        flags = (L.KRY_CSR_SYMMETRIC if symmetric else 0) | (L.KRY_CSR_BUILD_TRANSPOSE if build_transpose else 0)
        h = L.handle()
        call("kry_csr_create_coo", ctx._h, int(shape[0]), int(shape[1]), int(vals.shape[0]),
             _ptr(rows), _ptr(cols), _ptr(vals), flags, C.byref(h))
        return cls(ctx, h, symmetric)

    def combine(self, alpha=1.0, other=None, beta=1.0, diag=None, gamma=1.0, symmetric=None):
        """alpha*self [+ beta*other] [+ gamma*diag(d)] as a new operator in HBM (``kry_csr_combine``)."""
        if symmetric is None:
            symmetric = self.symmetric and (other is None or other.symmetric)
        d = None
        if diag is not None:
            d = _f64(diag)
            if d.shape != (self.shape[0],):
                raise ValueError("diagonal has the wrong size")
        h = L.handle()
        call("kry_csr_combine", self.ctx._h, self._h, float(alpha), other._h if other is not None else None,
             float(beta), _ptr(d) if d is not None else None, float(gamma),
             L.KRY_CSR_SYMMETRIC if symmetric else 0, C.byref(h))
        return DeviceCsr(self.ctx, h, symmetric)

    def to_dense(self):
        out = np.empty(self.shape, dtype=np.float64)
        call("kry_csr_to_dense", self._h, _ptr(out))
        return out

    @classmethod
    def poisson1d(cls, ctx, n, row_begin=0, row_end=-1):
        h = L.handle()
        call("kry_csr_create_poisson1d", ctx._h, int(n), int(row_begin), int(row_end), 0, C.byref(h))
        return cls(ctx, h, True)

    @classmethod
    def poisson2d(cls, ctx, g, row_begin=0, row_end=-1):
        h = L.handle()
        call("kry_csr_create_poisson2d", ctx._h, int(g), int(row_begin), int(row_end), 0, C.byref(h))
        return cls(ctx, h, True)

    @classmethod
    def convdiff3d(cls, ctx, m, gamma=0.5, row_begin=0, row_end=-1, build_transpose=False):
        h = L.handle()
        flags = L.KRY_CSR_BUILD_TRANSPOSE if build_transpose else 0
        call("kry_csr_create_convdiff3d", ctx._h, int(m), float(gamma), int(row_begin), int(row_end),
             flags, C.byref(h))
        return cls(ctx, h, False)

    # -- queries
    def download(self, transposed=False):
        nr = self.shape[1] if (transposed and not self.symmetric) else self.shape[0]
        indptr = np.empty(nr + 1, dtype=np.int32)
        indices = np.empty(self.nnz, dtype=np.int32)
        data = np.empty(self.nnz, dtype=np.float64)
        call("kry_csr_download", self._h, int(bool(transposed)), _ptr(indptr), _ptr(indices), _ptr(data))
        return indptr, indices, data

    def diagonal(self):
        d = np.empty(self.shape[0], dtype=np.float64)
        call("kry_csr_diagonal", self._h, _ptr(d))
        return d

    def build_transpose(self):
        call("kry_csr_build_transpose", self._h)

    def set_kernel(self, kind=L.KRY_SPMV_AUTO, tile_nnz=0, threads=0):
        call("kry_csr_set_kernel", self._h, int(kind), int(tile_nnz), int(threads))

    def shard_finalize(self, n_global, row_begin):
        call("kry_csr_shard_finalize", self._h, int(n_global), int(row_begin))
        nr, nc, nz = C.c_int64(), C.c_int64(), C.c_int64()
        call("kry_csr_shape", self._h, C.byref(nr), C.byref(nc), C.byref(nz))
        self.shape = (nr.value, nc.value)
        self.sharded = True
        self.n_global = int(n_global)

    def input_vector(self, init=None):
        """A vector usable as SpMV input (a shard's input carries the halo tail)."""
        n = self.shape[0] if self.sharded else self.shape[1]
        v = DeviceVector(self.ctx, n, capacity=self.shape[1])
        if init is not None:
            v.upload(init)
        return v

    # -- products
    def spmv(self, x, y, trans=False):
        call("kry_spmv", self._h, int(bool(trans)), x._h, y._h)
        return y

    def spmv_dot(self, x, y, dot_with, slot0=0, trans=False):
        n = len(dot_with)
        arr = (L.handle * max(n, 1))(*[(w._h if w is not None else None) for w in dot_with])
        call("kry_spmv_dot", self._h, int(bool(trans)), x._h, y._h, n, arr, int(slot0))
        return y

    def matvec(self, x, trans=False):
        """Host array in, host array out (LinearOperator.__mul__ bridge)."""
        nin = self.shape[0] if (trans or self.sharded) else self.shape[1]
        nout = self.shape[1] if trans else self.shape[0]
        x = _f64(x)
        if x.shape != (nin,):
            raise ValueError("input array size incompatible with operator dimensions")
        xv = DeviceVector(self.ctx, nin, capacity=max(nin, self.shape[1])).upload(x)
        yv = DeviceVector(self.ctx, nout)
        self.spmv(xv, yv, trans=trans)
        return yv.download()


def _axpby_array(ops):
    arr = (L.Axpby * max(len(ops), 1))()
    for k, o in enumerate(ops):
        arr[k].z = o["z"]._h
        arr[k].u = o["u"]._h if o.get("u") is not None else None
        arr[k].w = o["w"]._h if o.get("w") is not None else None
        arr[k].a = float(o.get("a", 1.0))
        arr[k].b = float(o.get("b", 1.0))
        arr[k].a_slot = int(o.get("a_slot", -1))
        arr[k].b_slot = int(o.get("b_slot", -1))
        arr[k].a_neg = int(o.get("a_neg", 0)) | (2 if o.get("a_div") else 0)
        arr[k].b_neg = int(o.get("b_neg", 0)) | (2 if o.get("b_div") else 0)
    return arr


def _dot_array(dots):
    darr = (L.DotSpec * max(len(dots), 1))()
    for k, (u, w) in enumerate(dots):
        darr[k].u = u._h
        darr[k].w = w._h
    return darr


def multi_axpy_dot(ctx, ops, dots=(), slot0=0, plane=None, phase=0):
    """ops: list of dicts(z, u, w, a, b, a_slot, b_slot, a_neg, b_neg); dots: list of (u, w).
    With `plane` (a ScalarPlane) the launch's finalize also runs phase `phase` of that plane's
    recurrence (kry_lls_multi_axpy_dot); the inner products then go to slots 0.."""
    if plane is not None:
        call("kry_lls_multi_axpy_dot", plane._h, int(phase), len(ops), _axpby_array(ops), len(dots), _dot_array(dots))
    else:
        call("kry_multi_axpy_dot", ctx._h, len(ops), _axpby_array(ops), len(dots), _dot_array(dots), int(slot0))


def spmv_axpby_dot(csr, x, op, dot=False, dot_with=None, slot0=0, trans=False, plane=None, phase=0):
    """z = a*(A x) + b*w in one launch (`op`: a dict like multi_axpy_dot's without `u`), optionally with
    the inner product dot_with . z (None: z . z) into slot `slot0`; with `plane` the finalize also runs
    phase `phase` of the plane's recurrence (kry_spmv_axpby_dot / kry_lls_spmv_axpby_dot)."""
    arr = _axpby_array([dict(op, u=None)])
    dw = dot_with._h if dot_with is not None else None
    if plane is not None:
        call("kry_lls_spmv_axpby_dot", plane._h, int(phase), csr._h, int(bool(trans)), x._h, arr, dw)
    else:
        call("kry_spmv_axpby_dot", csr._h, int(bool(trans)), x._h, arr, int(bool(dot) or dot_with is not None), dw, int(slot0))


class LaunchGraph(object):
    """A captured sequence of stand-alone launches (``kry_graph``): ``LaunchGraph.capture(ctx, fn)``
    records what ``fn()`` enqueues, ``launch(n)`` replays it n times.  Returns None where graphs are
    not available (host emulation)."""

    def __init__(self, ctx, handle):
        self.ctx, self._h = ctx, handle
        ctx._adopt(self)

    @classmethod
    def capture(cls, ctx, fn):
        h = L.handle()
        rc = L.lib.kry_graph_begin(ctx._h, C.byref(h))
        if rc in (L.KRY_ERR_UNSUPPORTED, L.KRY_ERR_STATE):      # emulation / per-launch profiling is on
            return None
        if rc != L.KRY_OK:
            raise L.KrylovDeviceError(rc, L.last_error())
        g = cls(ctx, h)
        try:
            fn()
        finally:
            rc = L.lib.kry_graph_end(h)
        if rc != L.KRY_OK:
            g._release()
            raise L.KrylovDeviceError(rc, L.last_error())
        return g

    replays = 0          # process-wide count of replayed sequences (read by the tests)

    def launch(self, times=1):
        call("kry_graph_launch", self._h, int(times))
        LaunchGraph.replays += int(times)

    def _release(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            L.lib.kry_graph_destroy(self._h)
            self._h = L.handle()

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass


class ScalarPlane(object):
    """Device-resident scalar plane of LSQR / LSMR / CRAIG / CRAIG-MR / SYMMLQ (``kry_lls``):
    named scalars in, ``step(phase)`` enqueues one phase of the recurrence as a launch of its own (the fused launches of
    ``multi_axpy_dot(plane=...)`` / ``spmv_axpby_dot(plane=...)`` run it in their finalize), the status block
    and the per-iteration history ring come back at the caller's check interval."""
    METHODS = {"lsqr": L.KRY_LLS_LSQR, "lsmr": L.KRY_LLS_LSMR, "craig": L.KRY_LLS_CRAIG,
               "craigmr": L.KRY_LLS_CRAIGMR, "symmlq": L.KRY_LLS_SYMMLQ}
    # coefficient slots of the context's scalar block (csrc/lls.cu)
    D0, D1, D2 = 0, 1, 2
    ALPHA, U_DIV, NV_A, NV_B, V_DIV = 8, 9, 10, 11, 12
    C0, C1, C2, C3, C4, C5, C6, C7 = 13, 14, 15, 16, 17, 18, 19, 20
    _names = {}

    def __init__(self, ctx, method):
        self.ctx = ctx
        self.method = method
        self._h = L.handle()
        call("kry_lls_create", ctx._h, self.METHODS[method], C.byref(self._h))
        if method not in self._names:
            names, i = [], 0
            while True:
                nm = L.lib.kry_lls_scalar_name(self.METHODS[method], i)
                if not nm:
                    break
                names.append(nm.decode())
                i += 1
            self._names[method] = names
        self.names = self._names[method]
        self._hist_read = 0
        ctx._adopt(self)

    def _release(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            L.lib.kry_lls_destroy(self._h)
            self._h = L.handle()

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def setup(self, scalars, window=5, istop=0, itn=0, nmatvec=0, itnlim=0, damp=0.0, atol=0.0, btol=0.0,
              ctol=0.0, etol=0.0, rtol=0.0, shift=0.0, eps=2.220446049250313e-16):
        unknown = set(scalars) - set(self.names)
        if unknown:
            raise KeyError("unknown %s scalars: %s" % (self.method, sorted(unknown)))
        p = L.LlsParams()
        p.window, p.istop, p.itn, p.nmatvec, p.itnlim = int(window), int(istop), int(itn), int(nmatvec), int(itnlim)
        p.damp, p.atol, p.btol, p.ctol = float(damp), float(atol), float(btol), float(ctol)
        p.etol, p.rtol, p.shift, p.eps = float(etol), float(rtol), float(shift), float(eps)
        vals = (C.c_double * len(self.names))(*[float(scalars.get(nm, 0.0)) for nm in self.names])
        call("kry_lls_setup", self._h, C.byref(p), vals, len(self.names))
        self._hist_read = 0

    def step(self, phase):
        call("kry_lls_step", self._h, int(phase))

    def status(self):
        """(status struct, dict of the method's scalars) -- one synchronising D2H."""
        st = L.LlsStatus()
        vals = (C.c_double * len(self.names))()
        call("kry_lls_status", self._h, C.byref(st), vals, len(self.names))
        return st, dict(zip(self.names, vals[:len(self.names)]))

    def drain_history(self, st):
        count = st.hist_count - self._hist_read
        if count <= 0:
            return np.empty((0, L.KRY_LLS_HIST_WIDTH))
        buf = np.empty((count, L.KRY_LLS_HIST_WIDTH), dtype=np.float64)
        call("kry_lls_history", self._h, self._hist_read, count, _ptr(buf))
        self._hist_read = st.hist_count
        return buf

    def release_gate(self):
        call("kry_lls_release_gate", self._h)


METHODS = {"cg": L.KRY_CG, "bicgstab": L.KRY_BICGSTAB, "cgs": L.KRY_CGS, "tfqmr": L.KRY_TFQMR,
           "minres": L.KRY_MINRES}


class DeviceSolver(object):
    """Device-resident Krylov iteration (``kry_solver``)."""

    def __init__(self, ctx, method, A):
        self.ctx = ctx
        self.A = A
        self.method = method
        self.n = A.shape[0]
        self._h = L.handle()
        call("kry_solver_create", ctx._h, METHODS[method], A._h, C.byref(self._h))
        self._hist_read = 0
        ctx._adopt(self)

    def _release(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            L.lib.kry_solver_destroy(self._h)
            self._h = L.handle()

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def set_precon_diag(self, diag, mode=1):
        if diag is None:
            call("kry_solver_set_precon_diag", self._h, None, 0)
        else:
            d = _f64(diag)
            if d.shape != (self.n,):
                raise ValueError("preconditioner diagonal has the wrong size")
            call("kry_solver_set_precon_diag", self._h, _ptr(d), int(mode))

    @staticmethod
    def params(abstol=1.0e-8, reltol=1.0e-6, matvec_max=0, check_curvature=True, guess_supplied=False,
               shift=0.0, rtol=1.0e-12, etol=1.0e-6, window=5):
        p = L.SolverParams()
        p.abstol, p.reltol, p.matvec_max = float(abstol), float(reltol), int(matvec_max)
        p.check_curvature, p.guess_supplied = int(bool(check_curvature)), int(bool(guess_supplied))
        p.shift, p.rtol, p.etol, p.window = float(shift), float(rtol), float(etol), int(window)
        return p

    def setup(self, rhs, guess=None, **kw):
        rhs = _f64(rhs)
        if rhs.shape != (self.n,):
            raise ValueError("right-hand side size incompatible with operator dimensions")
        g = None
        if guess is not None:
            g = _f64(guess)
            if g.shape != (self.n,):
                raise ValueError("initial guess size incompatible with operator dimensions")
        p = self.params(guess_supplied=guess is not None, **kw)
        call("kry_solver_setup", self._h, _ptr(rhs), _ptr(g) if g is not None else None, C.byref(p))
        self._hist_read = 0

    def setup_dev(self, rhs_vec, guess_vec=None, **kw):
        p = self.params(guess_supplied=guess_vec is not None, **kw)
        call("kry_solver_setup_dev", self._h, rhs_vec._h, guess_vec._h if guess_vec is not None else None,
             C.byref(p))
        self._hist_read = 0

    def iterate(self, n_iters):
        call("kry_solver_iterate", self._h, int(n_iters))

    def status(self):
        s = L.SolverStatus()
        call("kry_solver_status_read", self._h, C.byref(s))
        return s

    def status_enqueue(self, slot):
        """Queue a status snapshot behind everything enqueued so far (returns at once)."""
        call("kry_solver_status_enqueue", self._h, int(slot))

    def status_wait(self, slot):
        """Block until snapshot `slot` has landed -- not until later work has finished."""
        s = L.SolverStatus()
        call("kry_solver_status_wait", self._h, int(slot), C.byref(s))
        return s

    def drain_history(self, status=None, nowait=False):
        """New per-iteration entries since the last drain, shape (k, width).  With
        ``nowait`` the copy runs on the side stream (entries counted by a snapshot that has
        been waited for are final even while later iterations are still running)."""
        s = status if status is not None else self.status()
        count = s.hist_count - self._hist_read
        width = C.c_int32(0)
        fn = "kry_solver_history_nowait" if nowait else "kry_solver_history"
        if count <= 0:
            call(fn, self._h, 0, 0, None, C.byref(width))
            return np.empty((0, width.value))
        buf = np.empty((count, 2), dtype=np.float64)
        call(fn, self._h, self._hist_read, count, _ptr(buf), C.byref(width))
        self._hist_read = s.hist_count
        return buf.reshape(-1)[:count * width.value].reshape(count, width.value)

    def solution(self):
        x = result_pool.empty(self.n)
        call("kry_solver_solution", self._h, _ptr(x))
        return x

    def get_vector(self, name):
        v = result_pool.empty(self.n)
        call("kry_solver_get_vector", self._h, name.encode(), _ptr(v))
        return v

    def set_vector(self, name, a):
        a = _f64(a)
        if a.shape != (self.n,):
            raise ValueError("vector size incompatible with solver dimensions")
        call("kry_solver_set_vector", self._h, name.encode(), _ptr(a))

    def get_scalar(self, name):
        v = C.c_double(0.0)
        call("kry_solver_get_scalar", self._h, name.encode(), C.byref(v))
        return v.value

    def set_scalar(self, name, value):
        call("kry_solver_set_scalar", self._h, name.encode(), float(value))

    def run(self, check_interval=64):
        """Iterate until the device latches `done`; returns the final status."""
        while True:
            self.iterate(check_interval)
            s = self.status()
            if s.done:
                return s
