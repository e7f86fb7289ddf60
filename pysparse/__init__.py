"""Minimal stand-in for the (un-vendored, un-pinned, absent) Pysparse package.

Only what the reference's acceptance script needs (examples/bmark.py:10-11,34-43;
examples/demo_common.py:4-5,15-16): ``spmatrix.ll_mat_from_mtx`` and
``sparse.pysparseMatrix.PysparseMatrix`` with ``shape``, ``A*x``, ``takeDiagonal()``
and ``issym``.  Matrices export CSR so that ``PysparseLinearOperator`` can place
them in HBM; products asked of the matrix object itself (``rhs = A*e``) run on
the device as well.
"""
from . import spmatrix      # noqa: F401
