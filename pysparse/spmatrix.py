"""``pysparse.spmatrix`` stand-in: linked-list matrices are held as CSR."""
import numpy as np


class ll_mat(object):
    """Read-only sparse matrix in CSR clothing (enough for the demos)."""

    def __init__(self, shape, indptr, indices, data, issym=False):
        self.shape = (int(shape[0]), int(shape[1]))
        self.indptr, self.indices, self.data = indptr, indices, data
        self.issym = bool(issym)
        self.nnz = int(len(data))
        self._op = None

    def to_csr_arrays(self):
        return self.indptr, self.indices, self.data

    def _device(self):
        if self._op is None:
            from pykrylov_b200.linop import csr_operator
            self._op = csr_operator(self.shape, self.indptr, self.indices, self.data,
                                    symmetric=self.issym)
        return self._op

    def matvec(self, x, y):
        y[:] = self._device() * np.asarray(x, dtype=np.float64)

    def matvec_transp(self, x, y):
        y[:] = self._device().T * np.asarray(x, dtype=np.float64)

    def take_diagonal(self):
        return self._device().diagonal()


def ll_mat_from_mtx(path):
    """Matrix Market file -> matrix (symmetric storage is expanded; ``issym`` kept)."""
    from pykrylov_b200.mmio import read_mtx
    shape, indptr, indices, data, sym = read_mtx(path)
    return ll_mat(shape, indptr, indices, data, issym=sym)
