"""CPU tests of the host side: the LinearOperator protocol (modelled on the
reference's pykrylov/linop/tests/test_linop.py strategy), the Matrix Market
reader, the C-ABI library (loads, exports every declared symbol, fails loudly
without a GPU) and the package aliases.  No compute call reaches a GPU here."""
import ctypes
import os
import re

import numpy as np
import pytest
import scipy.io as sio
import scipy.sparse as sp

from conftest import ROOT, mtx
import pykrylov_b200.linop as lo
from pykrylov_b200.linop import ShapeError
from pykrylov_b200.tools.types import allowed_types
from pykrylov_b200.tools import check_symmetric, check_positive_definite, roots_quadratic, machine_epsilon


def dense_op(A, with_transp=True, with_adj=False, **kw):
    return lo.LinearOperator(A.shape[1], A.shape[0], lambda x: np.dot(A, x),
                             matvec_transp=(lambda x: np.dot(A.T, x)) if with_transp else None,
                             matvec_adj=(lambda x: np.dot(A.T.conjugate(), x)) if with_adj else None, **kw)


A23 = np.array([[1, 2, 3], [4, 5, 6]])
C22 = np.array([[1, 2], [3, 4]])
D23 = np.array([[1 + 1j, 2 - 2j, 3 + 1j], [4 - 2j, 5 + 3j, 6]])


# ------------------------------------------------------------------ protocol
def test_transpose_adjoint_inference_real():
    A = dense_op(A23, with_transp=False)
    assert A.T is None and A.H is None and hasattr(A, "_matvec") and hasattr(A, "dtype")
    A = dense_op(A23)
    assert A.T is not None and A.H is A.T
    assert A.T.T is A and A.H.H is A                      # linop.py:148-170, doc linop.rst:23-29
    assert A.shape == (2, 3) and A.T.shape == (3, 2)
    u, v = np.array([1, 1]), np.array([1, 1, 1])
    assert np.array_equal(A * v, [6, 15]) and np.array_equal(A._matvec(v), [6, 15])
    assert np.array_equal(A.T * u, [5, 7, 9]) and np.array_equal(A.H * u, [5, 7, 9])
    S = dense_op(np.array([[2., 1.], [1., 3.]]), with_transp=False, symmetric=True)
    assert S.T is S and S.H is S


def test_transpose_adjoint_inference_complex():
    rng = np.random.default_rng(0)
    D = dense_op(D23, with_adj=True, dtype=D23.dtype)
    assert D.T is not None and D.H is not None
    E = lo.LinearOperator(2, 3, lambda x: np.dot(D23.T, x), matvec_transp=lambda x: np.dot(D23, x),
                          dtype=D23.dtype)                 # E = D.T, adjoint must be inferred
    x = rng.random(2) + 1j * rng.random(2)
    y = rng.random(3) + 1j * rng.random(3)
    assert np.allclose(E * x, D.T * x)
    assert E.H is not None and np.allclose(E.H * y, np.dot(D23.conjugate(), y))
    assert np.allclose(E.H * y, E.rmatvec(y))
    F = lo.LinearOperator(2, 3, lambda x: np.dot(D23.T.conjugate(), x), dtype=D23.dtype)
    assert np.allclose(F * x, D.H * x)
    G = np.array([[1, 2 - 2j], [2 + 2j, 4]])
    Gop = dense_op(G, with_adj=True, dtype=G.dtype)
    assert np.allclose(Gop * x, Gop.H * x) and np.allclose(Gop.T.H * x, Gop.H.T * x)
    assert np.allclose(D.bar * y, np.dot(D23.conjugate(), y)) and D.bar.bar is D
    assert D.T.T is D and D.H.H is D


def test_algebra_and_error_types():
    A, B, C = dense_op(A23), dense_op(A23.T.copy()), dense_op(C22)
    u, v = np.array([1, 1]), np.array([1, 1, 1])
    assert np.array_equal((A * 2) * v, A * (2 * v)) and np.array_equal((2 * A) * v, (A * 2) * v)
    assert np.array_equal((A / 2) * v, A * (v / 2)) and np.array_equal((-A) * v, A * (-v))
    assert np.array_equal((A - A) * v, [0, 0]) and np.array_equal((A + A) * v, [12, 30])
    assert np.array_equal((C ** 2) * u, [17, 37]) and np.array_equal((C * C) * u, [17, 37])
    assert np.array_equal((A * B) * u, np.dot(A23, np.dot(A23.T, u)))
    assert np.array_equal((A * B).T * u, np.dot(A23, np.dot(A23.T, u)))
    for op in (A + A, A - A, -A, 2 * A, A * 2, A / 2, C ** 2):
        assert isinstance(op, lo.LinearOperator)
    assert isinstance(A * 0, lo.ZeroOperator) and isinstance(C ** 0, lo.IdentityOperator)
    for bad in (3, v):
        with pytest.raises(ValueError):
            A + bad
        with pytest.raises(ValueError):
            A - bad
    with pytest.raises(ShapeError):
        A + B
    with pytest.raises(ShapeError):
        A - B
    with pytest.raises(ValueError):
        A * u                                               # wrong size: linop.py:283-296
    with pytest.raises(ShapeError):
        A * A
    with pytest.raises(ValueError):
        A * [1, 1, 1]                                       # not an ndarray: linop.py:369
    with pytest.raises(ValueError):
        A / B
    with pytest.raises(ValueError):
        A / u
    with pytest.raises(ZeroDivisionError):
        A / 0
    with pytest.raises(ShapeError):
        A ** 2
    with pytest.raises(ValueError):
        C ** -2
    with pytest.raises(ValueError):
        C ** 2.1


def test_dtype_promotion_over_all_allowed_types():
    for dt_op in allowed_types:
        for dt_in in allowed_types:
            out = np.result_type(dt_op, dt_in)
            x = np.array([1, 1, 1]).astype(dt_in)
            assert (dense_op(A23, dtype=dt_op) * x).dtype == out
            assert (lo.IdentityOperator(3, dtype=dt_op) * x).dtype == out
            assert (lo.ZeroOperator(3, 2, dtype=dt_op) * x).dtype == out
    A = dense_op(A23)
    with pytest.raises(TypeError):
        A.dtype = "nope"
    A.dtype = np.float32
    assert A.dtype == np.float32


def test_counters_and_call_alias():
    A = dense_op(A23)
    v = np.ones(3)
    A * v
    A(v)
    assert A.nMatvec == 2
    A.reset_counters()
    assert A.nMatvec == 0
    assert "Unsymmetric" in repr(A) and "(2,3)" in repr(A)


def test_identity_diagonal_zero():
    I3 = lo.IdentityOperator(3)
    x = np.array([1., 2., 3.])
    assert np.array_equal(I3 * x, x) and I3.T is I3 and I3.H is I3
    d = np.array([1., -2., 3.])
    Dg = lo.DiagonalOperator(d)
    assert np.array_equal(Dg * x, d * x) and Dg.T is Dg and Dg.H is Dg and Dg.symmetric
    assert np.array_equal(Dg.diag, d) and np.array_equal(abs(Dg).diag, np.abs(d))
    with pytest.raises(ValueError):
        lo.sqrt(Dg)
    assert np.allclose(lo.sqrt(abs(Dg)).diag ** 2, np.abs(d))
    with pytest.raises(ValueError):
        lo.DiagonalOperator(np.ones((2, 2)))
    dc = np.array([1 + 1j, 2.0, 3 - 1j])
    Dc = lo.DiagonalOperator(dc)
    assert not Dc.hermitian and np.allclose(Dc.H * x, dc.conjugate() * x)
    Z = lo.ZeroOperator(3, 2)
    assert np.array_equal(Z * x, [0, 0]) and np.array_equal(Z.T * np.ones(2), [0, 0, 0])
    with pytest.raises(ValueError):
        Z * np.ones(2)
    assert abs(Z) is Z and lo.sqrt(Z) is Z and lo.ZeroOperator(2, 2).symmetric


def test_reduced_operators():
    M = np.arange(20, dtype=float).reshape(4, 5)
    op = dense_op(M)
    rows, cols = [0, 2], [1, 3, 4]
    R = lo.ReducedLinearOperator(op, rows, cols)
    assert R.shape == (2, 3)
    x, y = np.array([1., 2., 3.]), np.array([1., -1.])
    assert np.array_equal(R * x, M[np.ix_(rows, cols)] @ x)
    assert np.array_equal(R.T * y, M[np.ix_(rows, cols)].T @ y)
    Ssym = M[:4, :4] + M[:4, :4].T
    sop = dense_op(Ssym, symmetric=True)
    idx = [0, 3]
    SR = lo.SymmetricallyReducedLinearOperator(sop, idx)
    assert SR.symmetric and np.array_equal(SR * y, Ssym[np.ix_(idx, idx)] @ y)


def test_linop_from_ndarray_and_to_array():
    M = np.array([[1., 2.], [3., 4.], [5., 6.]])
    op = lo.linop_from_ndarray(M)
    assert np.array_equal(op.to_array(), M) and np.array_equal(op.full(), M)
    assert np.array_equal(op.T.to_array(), M.T)
    with pytest.raises(ValueError):
        lo.linop_from_ndarray(M, symmetric=True, hermitian=False)
    Hm = np.array([[1, 2 - 2j], [2 + 2j, 4]])
    hop = lo.linop_from_ndarray(Hm, hermitian=True)
    assert hop.H is hop and np.allclose(hop.T.to_array(), Hm.T)


def test_coord_operator_complex_stays_on_host():
    vals = np.array([1 + 1j, 2.0, 3 - 1j])
    rows, cols = np.array([0, 1, 1]), np.array([0, 0, 1])
    op = lo.CoordLinearOperator(vals, rows, cols, nargin=2, nargout=2)
    x = np.array([1.0, 2.0])
    assert np.allclose(op * x, [1 + 1j, 2 + 2 * (3 - 1j)])
    assert np.allclose(op.T * x, [(1 + 1j) + 4.0, 2 * (3 - 1j)])
    with pytest.raises(IndexError):
        lo.CoordLinearOperator(vals, rows, cols)            # reference default sizes: max() w/o +1


def test_coo_to_csr_keeps_reference_accumulation_order():
    rng = np.random.default_rng(3)
    n, k = 9, 40
    rows, cols = rng.integers(0, n, k), rng.integers(0, n, k)
    lower = rows >= cols
    rows, cols = rows[lower], cols[lower]
    vals = rng.standard_normal(len(rows))
    ip, c, v = lo.linop._coo_to_csr_in_arrival_order(vals, rows, cols, n, True)
    x = rng.standard_normal(n)
    y = np.zeros(n)                                         # the reference's loop, linop.py:657-664
    for q in range(len(vals)):
        y[rows[q]] += vals[q] * x[cols[q]]
        if rows[q] != cols[q]:
            y[cols[q]] += vals[q] * x[rows[q]]
    z = np.zeros(n)
    for i in range(n):
        s = 0.0
        for q in range(ip[i], ip[i + 1]):
            s += v[q] * x[c[q]]
        z[i] = s
    assert np.array_equal(y, z)


def test_tools():
    S = dense_op(np.array([[2., 1.], [1., 3.]]), symmetric=True)
    N = dense_op(np.array([[2., 1.], [0., 3.]]))
    assert check_symmetric(S) and not check_symmetric(N) and not check_symmetric(dense_op(A23))
    assert check_positive_definite(S) and not check_positive_definite(-S)
    assert machine_epsilon() == np.finfo(float).eps
    assert np.allclose(sorted(roots_quadratic(1, -3, 2)), [1, 2])
    assert roots_quadratic(0, 0, 1) == [] and roots_quadratic(0, 0, 0) == [0.0]
    assert roots_quadratic(1, 0, 1) == [] and np.allclose(roots_quadratic(0, 2, -4), [2.0])


# ------------------------------------------------------------------ data formats
@pytest.mark.parametrize("name", ["1138bus", "jpwh_991", "GD97_b"])
def test_mtx_reader_equals_scipy(name, golden):
    from pykrylov_b200.mmio import read_mtx
    import zlib
    shape, ip, ix, dv, sym = read_mtx(mtx(name))
    M = sp.csr_matrix(sio.mmread(mtx(name)))
    M.sort_indices()
    assert shape == M.shape and sym == (name != "jpwh_991")
    assert ip.dtype == np.int32 and ix.dtype == np.int32 and dv.dtype == np.float64
    assert np.array_equal(ip, M.indptr) and np.array_equal(ix, M.indices) and np.array_equal(dv, M.data)
    assert zlib.crc32(ix.tobytes()) == golden["csr/" + name]["indices_crc"]


def test_mtx_reader_edge_cases(tmp_path):
    from pykrylov_b200.mmio import read_mtx, MatrixMarketError
    p = tmp_path / "a.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real general\n% c\n3 4 4\n3 4 1.5\n1 1 2\n1 1 3\n2 3 -1\n")
    shape, ip, ix, dv, sym = read_mtx(str(p))
    assert shape == (3, 4) and list(ip) == [0, 1, 2, 3] and list(ix) == [0, 2, 3] and list(dv) == [5.0, -1.0, 1.5]
    p.write_text("%%MatrixMarket matrix coordinate pattern skew-symmetric\n2 2 1\n2 1\n")
    shape, ip, ix, dv, sym = read_mtx(str(p))
    assert list(ix) == [1, 0] and list(dv) == [-1.0, 1.0] and not sym
    p.write_text("%%MatrixMarket matrix coordinate real general\n2 2 0\n")
    shape, ip, ix, dv, sym = read_mtx(str(p))
    assert list(ip) == [0, 0, 0] and len(dv) == 0
    p.write_text("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n")
    with pytest.raises(MatrixMarketError):
        read_mtx(str(p))
    p.write_text("hello\n")
    with pytest.raises(MatrixMarketError):
        read_mtx(str(p))


def test_gallery_host_definitions_match_csr():
    from pykrylov_b200.gallery import Poisson1dMatvec, Poisson2dMatvec, poisson2d_csr_arrays
    from oracle import krylov_ref as kr
    x = np.random.default_rng(2).standard_normal(49)
    assert np.array_equal(Poisson2dMatvec(x.copy()), kr.poisson2d_matvec(x.copy()))
    assert np.array_equal(Poisson1dMatvec(x.copy()), kr.poisson1d_matvec(x.copy()))
    ip, ix, dv = poisson2d_csr_arrays(7)
    rp, ri, rd = kr.poisson2d_csr(7)
    assert np.array_equal(ip, rp) and np.array_equal(ix, ri) and np.array_equal(dv, rd)


# ------------------------------------------------------------------ C ABI
def declared_symbols():
    text = open(os.path.join(ROOT, "include", "krylov_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kry_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from pykrylov_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 50
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(raw, name), "libkrylov_b200.so does not export %s" % name
        assert name in _lib.PROTOTYPES, "no ctypes prototype for %s" % name
    assert set(_lib.PROTOTYPES) <= set(names)
    assert _lib.lib.kry_abi_version() == 1


def test_struct_layouts_match_header():
    from pykrylov_b200 import _lib
    assert ctypes.sizeof(_lib.SolverParams) == 64
    assert ctypes.sizeof(_lib.SolverStatus) == 16 + 24 + 24 + 128
    assert ctypes.sizeof(_lib.Axpby) == 24 + 16 + 16 and ctypes.sizeof(_lib.DotSpec) == 16


def test_no_gpu_means_loud_failure_not_cpu_fallback():
    from pykrylov_b200 import _lib
    from pykrylov_b200.device import Context, device_count
    try:
        n = device_count()
    except _lib.KrylovDeviceError:
        n = 0
    if n > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.KrylovDeviceError) as ei:
        Context(0)
    assert ei.value.status == _lib.KRY_ERR_CUDA
    with pytest.raises(_lib.KrylovDeviceError):
        lo.csr_operator((2, 2), [0, 1, 2], [0, 1], [1.0, 1.0])
    # NULL handles are rejected by the ABI itself
    assert _lib.lib.kry_ctx_sync(None) == _lib.KRY_ERR_INVALID
    assert "NULL" in _lib.last_error()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pykrylov_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "refpykrylov" not in text and "krylov_ref" not in text, f
                # ... nor the host emulation of tests/emu: the product neither loads libkrylov_emu
                # nor defines KRY_EMULATE (the guards in csrc/*.cuh only *test* the macro)
                assert "libkrylov_emu" not in text and "build_emu" not in text, f
                assert not re.search(r"#\s*include[^\n]*emu_", text), f
                assert not re.search(r"#\s*define\s+KRY_EMULATE", text), f
    make = open(os.path.join(pkg, "csrc", "Makefile")).read()
    assert "KRY_EMULATE" not in make


def test_pykrylov_alias_and_pysparse_shim_import():
    import pykrylov
    from pykrylov.linop import PysparseLinearOperator
    from pykrylov.cgs import CGS
    from pykrylov.tfqmr import TFQMR
    from pykrylov.bicgstab import BiCGSTAB
    from pykrylov.cg import CG
    from pykrylov.minres import Minres
    import pykrylov_b200.cgs
    assert CGS is pykrylov_b200.cgs.CGS and PysparseLinearOperator is lo.PysparseLinearOperator
    assert {k.__name__ for k in (CG, CGS, TFQMR, BiCGSTAB, Minres)} == {"CG", "CGS", "TFQMR", "BiCGSTAB", "Minres"}
    from pysparse import spmatrix
    from pysparse.sparse.pysparseMatrix import PysparseMatrix
    A = PysparseMatrix(matrix=spmatrix.ll_mat_from_mtx(mtx("jpwh_991")))
    assert A.shape == (991, 991) and not A.issym and len(A.to_csr_arrays()[2]) == 6027


def test_krylov_method_contract():
    from pykrylov_b200.generic import KrylovMethod
    k = KrylovMethod("op", matvec_max=7, outputStream=None, reltol=1e-3)   # unknown kwargs ignored
    assert (k.abstol, k.reltol, k.precon, k.nMatvec, k.converged, k.bestSolution) == (1e-8, 1e-3, None, 0, False, None)
    assert k.residHistory == [] and k.x is None
    with pytest.raises(NotImplementedError):
        k.solve(np.ones(2))


def test_pinned_result_pool_recycles_blocks_and_respects_caps():
    """device.PinnedPool hands out NumPy arrays backed by page-locked blocks and takes the
    block back when the last view dies (allocator injected: no GPU needed here)."""
    import ctypes as C
    import gc
    from pykrylov_b200.device import PinnedPool
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    libc.malloc.argtypes = [C.c_size_t]
    libc.free.argtypes = [C.c_void_p]
    freed = []
    pool = PinnedPool(alloc=lambda nb: libc.malloc(nb), free=lambda p: (freed.append(p), libc.free(p)),
                      min_bytes=1024, live_cap=10 << 20, idle_cap=3 << 20)
    a = pool.empty(1000)
    assert a.shape == (1000,) and a.dtype == np.float64 and a.flags.writeable and a.flags.c_contiguous
    assert pool.live_bytes == 1 << 20 and pool.misses == 1
    a[:] = 3.0
    view = a[10:20]
    del a
    gc.collect()
    assert pool.live_bytes == 1 << 20 and view.sum() == 30.0       # a view keeps the block alive
    del view
    gc.collect()
    assert pool.live_bytes == 0 and pool.idle_bytes == 1 << 20 and not freed
    b = pool.empty(1000)                                            # same size class: recycled
    assert pool.hits == 1 and pool.idle_bytes == 0
    assert pool.empty(10).base is None                              # below min_bytes: ordinary array
    many = [pool.empty(300000) for _ in range(5)]                   # 3 MiB each; live cap 10 MiB
    assert [m.base is not None for m in many] == [True, True, True, False, False]
    del many, b
    gc.collect()
    assert pool.live_bytes == 0 and pool.idle_bytes <= 3 << 20 and len(freed) >= 1   # idle cap: rest unpinned
    # astype(copy=False) -- what the solvers do with the solution -- keeps the block
    x = pool.empty(2000).astype(np.float64, copy=False)
    assert x.base is not None
    del x
    gc.collect()
    pool.trim()
    assert pool.idle_bytes == 0
    # a failing allocator (no GPU / out of pinned memory) falls back to pageable arrays
    none = PinnedPool(alloc=lambda nb: None, free=lambda p: None, min_bytes=8)
    assert none.empty(100).base is None and none.live_bytes == 0
