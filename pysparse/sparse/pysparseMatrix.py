"""``pysparse.sparse.pysparseMatrix`` stand-in."""
import numpy as np


class PysparseMatrix(object):
    """Wrapper with the few methods the reference's scripts call."""

    def __init__(self, matrix=None, **kwargs):
        if matrix is None:
            raise ValueError("this stand-in only wraps an existing matrix (matrix=...)")
        self.matrix = matrix
        self.shape = matrix.shape
        self.issym = matrix.issym

    def isSymmetric(self):
        return self.issym

    def getShape(self):
        return self.shape

    def to_csr_arrays(self):
        return self.matrix.to_csr_arrays()

    def takeDiagonal(self):
        return self.matrix.take_diagonal()

    def __mul__(self, x):
        x = np.asarray(x, dtype=np.float64)
        y = np.empty(self.shape[0])
        self.matrix.matvec(x, y)
        return y

    def __rmul__(self, x):
        x = np.asarray(x, dtype=np.float64)
        y = np.empty(self.shape[1])
        self.matrix.matvec_transp(x, y)
        return y
