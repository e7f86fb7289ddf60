"""Matrix Market (coordinate) reader -> CSR arrays, NumPy only.

The reference loads its fixtures through Pysparse's ``ll_mat_from_mtx``
(examples/bmark.py:34); Pysparse is not available, so this reader produces the
CSR the device operator needs: symmetric / skew-symmetric storage is expanded,
duplicates are summed, column indices are sorted inside each row (the layout
scipy.sparse produces -- pinned bit-for-bit in tests/test_host.py).
"""
import numpy as np


class MatrixMarketError(ValueError):
    pass


def read_mtx(path):
    """Returns (shape, indptr[int32], indices[int32], data[float64], symmetric_flag)."""
    with open(path, "rb") as fh:
        header = fh.readline().decode("ascii", "replace").strip().split()
        if len(header) < 5 or header[0] != "%%MatrixMarket" or header[1].lower() != "matrix":
            raise MatrixMarketError("not a Matrix Market matrix file: %s" % path)
        fmt, field, symm = header[2].lower(), header[3].lower(), header[4].lower()
        if fmt != "coordinate":
            raise MatrixMarketError("only coordinate format is supported (got %s)" % fmt)
        if field not in ("real", "integer", "pattern"):
            raise MatrixMarketError("unsupported field %s" % field)
        if symm not in ("general", "symmetric", "skew-symmetric"):
            raise MatrixMarketError("unsupported symmetry %s" % symm)
        line = fh.readline()
        while line and (line.lstrip().startswith(b"%") or not line.strip()):
            line = fh.readline()
        nrow, ncol, nnz = (int(tok) for tok in line.split()[:3])
        body = np.loadtxt(fh, ndmin=2) if nnz > 0 else np.zeros((0, 3))
    if body.shape[0] != nnz:
        raise MatrixMarketError("expected %d entries, found %d" % (nnz, body.shape[0]))
    rows = body[:, 0].astype(np.int64) - 1
    cols = body[:, 1].astype(np.int64) - 1
    vals = np.ones(nnz) if field == "pattern" else body[:, 2].astype(np.float64)
    if symm != "general":
        off = rows != cols
        sign = -1.0 if symm == "skew-symmetric" else 1.0
        rows, cols, vals = (np.concatenate([rows, cols[off]]), np.concatenate([cols, rows[off]]),
                            np.concatenate([vals, sign * vals[off]]))
    indptr, indices, data = coo_to_csr(nrow, rows, cols, vals)
    return (nrow, ncol), indptr, indices, data, symm == "symmetric"


def coo_to_csr(nrow, rows, cols, vals):
    """Sorted-column CSR with duplicates summed (in their order of appearance)."""
    order = np.lexsort((cols, rows))             # stable: row major, then column
    rows, cols, vals = rows[order], cols[order], vals[order]
    if len(vals):
        new = np.ones(len(vals), dtype=bool)
        new[1:] = (rows[1:] != rows[:-1]) | (cols[1:] != cols[:-1])
        if not new.all():
            group = np.cumsum(new) - 1
            summed = np.zeros(group[-1] + 1)
            np.add.at(summed, group, vals)
            rows, cols, vals = rows[new], cols[new], summed
    counts = np.bincount(rows, minlength=nrow) if len(vals) else np.zeros(nrow, dtype=np.int64)
    indptr = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    return indptr, cols.astype(np.int32), np.ascontiguousarray(vals, dtype=np.float64)
