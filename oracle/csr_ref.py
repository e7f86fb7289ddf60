"""ctypes front-end of oracle/csr_ref.c + fixture loading.  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libcsr_ref.so")


def build():
    """Compile the C restatement (gcc only; seconds)."""
    src = os.path.join(_HERE, "csr_ref.c")
    if os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(src):
        return _SO
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-shared",
                           "-o", _SO, src])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        p = C.c_void_p
        _lib.csr_matvec_ref.argtypes = [C.c_int64, p, p, p, p, p]
        _lib.csr_matvec_ref.restype = None
        _lib.csr_matvec_transp_ref.argtypes = [C.c_int64, C.c_int64, p, p, p, p, p]
        _lib.csr_matvec_transp_ref.restype = None
        _lib.dot_seq_ref.argtypes = [C.c_int64, p, p]
        _lib.dot_seq_ref.restype = C.c_double
    return _lib


class CsrRef(object):
    """CSR matrix with the oracle's matvec / transposed matvec."""

    def __init__(self, shape, indptr, indices, data):
        self.shape = (int(shape[0]), int(shape[1]))
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int32)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32)
        self.data = np.ascontiguousarray(data, dtype=np.float64)
        self.nnz = int(self.data.shape[0])

    @classmethod
    def from_scipy(cls, M):
        M = M.tocsr()
        M.sort_indices()
        return cls(M.shape, M.indptr, M.indices, M.data)

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.data, self.indices, self.indptr), shape=self.shape)

    def matvec(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.shape == (self.shape[1],)
        y = np.empty(self.shape[0])
        lib().csr_matvec_ref(self.shape[0], self.indptr.ctypes.data, self.indices.ctypes.data,
                             self.data.ctypes.data, x.ctypes.data, y.ctypes.data)
        return y

    def rmatvec(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        assert x.shape == (self.shape[0],)
        y = np.empty(self.shape[1])
        lib().csr_matvec_transp_ref(self.shape[0], self.shape[1], self.indptr.ctypes.data,
                                    self.indices.ctypes.data, self.data.ctypes.data, x.ctypes.data,
                                    y.ctypes.data)
        return y

    __call__ = matvec


def load_mtx(path):
    """Matrix Market -> CsrRef exactly as the survey's stand-in operator is built:
    scipy.io.mmread (expands symmetric storage) -> CSR -> sorted indices."""
    import scipy.io as sio
    import scipy.sparse as sp
    return CsrRef.from_scipy(sp.csr_matrix(sio.mmread(path)))
