"""Problem gallery.

``Poisson1dMatvec`` / ``Poisson2dMatvec`` keep the reference's matrix-free
definitions (pykrylov/gallery/gallery.py:3-29) as host-side problem statements;
``poisson1d_operator`` / ``poisson2d_operator`` / ``convdiff3d_operator`` build the
same operators as CSR directly in HBM (libkrylov_b200's device gallery) and are
what the GPU hot path iterates on.
"""
from math import sqrt

import numpy as np


def Poisson1dMatvec(x):
    """y = T x with T = tridiag(-1, 2, -1)."""
    y = 2 * x
    y[:-1] -= x[1:]
    y[1:] -= x[:-1]
    return y


def Poisson2dMatvec(x):
    """y = A x, A the 5-point Laplacian on an n x n grid (n = sqrt(len(x))):
    diagonal 4, the four grid neighbours -1, homogeneous Dirichlet boundary."""
    n = int(sqrt(x.shape[0]))
    X = x[: n * n].reshape(n, n)
    y = 4 * x
    Y = y[: n * n].reshape(n, n)
    Y[1:, :] -= X[:-1, :]          # block sub-diagonal
    Y[:-1, :] -= X[1:, :]          # block super-diagonal
    Y[:-1, :-1] -= X[:-1, 1:]      # within blocks 0..n-2: right neighbour, then left
    Y[:-1, 1:] -= X[:-1, :-1]
    Y[-1, 1:] -= X[-1, :-1]        # last block: left neighbour, then right
    Y[-1, :-1] -= X[-1, 1:]
    return y


def poisson1d_operator(n, context=None):
    from ..linop import CsrLinearOperator
    from ..device import DeviceCsr, default_context
    ctx = context or default_context()
    return CsrLinearOperator(DeviceCsr.poisson1d(ctx, n))


def poisson2d_operator(grid, context=None):
    from ..linop import CsrLinearOperator
    from ..device import DeviceCsr, default_context
    ctx = context or default_context()
    return CsrLinearOperator(DeviceCsr.poisson2d(ctx, grid))


def convdiff3d_operator(grid, gamma=0.5, context=None, build_transpose=False):
    from ..linop import CsrLinearOperator
    from ..device import DeviceCsr, default_context
    ctx = context or default_context()
    return CsrLinearOperator(DeviceCsr.convdiff3d(ctx, grid, gamma, build_transpose=build_transpose))


def poisson2d_csr_arrays(grid):
    """(indptr, indices, data) of the grid x grid 5-point Laplacian, host side."""
    g = int(grid)
    n = g * g
    i = np.arange(n, dtype=np.int64)
    cx = i % g
    keep = np.stack([i >= g, cx > 0, np.ones(n, bool), cx < g - 1, i < n - g], axis=1)
    cols = np.stack([i - g, i - 1, i, i + 1, i + g], axis=1)
    vals = np.broadcast_to(np.array([-1.0, -1.0, 4.0, -1.0, -1.0]), cols.shape)
    indptr = np.concatenate([[0], np.cumsum(keep.sum(axis=1))]).astype(np.int32)
    return indptr, cols[keep].astype(np.int32), np.ascontiguousarray(vals[keep])


def kron_sym_jpwh(mtx_path, k=1009):
    """BASELINE config 3: A = kron(I_k, S) with S = (B + B^T)/2 and B the
    Matrix Market file at ``mtx_path`` (examples/jpwh_991.mtx).  Returns host CSR arrays
    ((n, n), indptr[int32], indices[int32], data) with sorted columns; k = 1009 gives
    n = 999 919, nnz = 6 404 123."""
    from ..mmio import read_mtx, coo_to_csr
    (m, _), ip, ix, dv, _ = read_mtx(mtx_path)
    rows = np.repeat(np.arange(m, dtype=np.int64), np.diff(ip))
    cols = ix.astype(np.int64)
    # (B + B^T) * 0.5: coincident entries are summed first, then halved (scipy's expression order)
    sip, six, sdv = coo_to_csr(m, np.concatenate([rows, cols]), np.concatenate([cols, rows]),
                               np.concatenate([dv, dv]))
    sdv = sdv * 0.5
    nnz1 = len(sdv)
    k = int(k)
    indptr = (np.arange(k, dtype=np.int64)[:, None] * nnz1 + sip[None, :-1].astype(np.int64)).reshape(-1)
    indptr = np.concatenate([indptr, [k * nnz1]]).astype(np.int32)
    indices = (np.arange(k, dtype=np.int64)[:, None] * m + six[None, :].astype(np.int64)).reshape(-1).astype(np.int32)
    return (k * m, k * m), indptr, indices, np.tile(sdv, k)
