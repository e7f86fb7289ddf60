"""GPU parity tests (run with `-m gpu` on a B200): every CUDA path is called through
the C ABI (ctypes) and compared with the oracle on identical, seeded inputs.

Bars (BASELINE.json north_star):
  * integer / indexing work and the SpMV row sums: BIT-EXACT;
  * floating point per iteration from an identical state: relative <= 1e-12;
  * final residual: agreement <= 1e-8.
"""
import os
import zlib

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import mtx
from oracle import krylov_ref as kr
from oracle.csr_ref import CsrRef, load_mtx

pytestmark = pytest.mark.gpu

RTOL_STEP = 1.0e-12      # per-iteration floating point bar
MARGINS = {}             # label -> largest per-step relative error seen in this session (conftest dumps it)


def within(err, tol, label):
    """Assert err <= tol and remember the largest error seen per label: the session writes them to
    gpurun_out/parity_margins.json, which is where the bars quoted in DESIGN.md section 3 come from."""
    err = float(err)
    cur = MARGINS.get(label)
    if cur is None or err > cur["max_err"]:
        MARGINS[label] = {"max_err": err, "bar": float(tol)}
    assert err <= tol, (label, err, tol)
RTOL_FINAL = 1.0e-8      # final residual bar


def L():
    from pykrylov_b200 import _lib
    return _lib


def dev():
    from pykrylov_b200 import device
    return device


def rel(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    scale = np.max(np.abs(b)) if b.size else 0.0
    if scale == 0.0:
        return float(np.max(np.abs(a))) if a.size else 0.0
    return float(np.max(np.abs(a - b)) / scale)


def srel(a, b):
    return abs(a - b) / abs(b) if b != 0 else abs(a)


def upload(ctx, M, symmetric=False, transpose=False):
    return dev().DeviceCsr.from_arrays(ctx, M.shape, M.indptr, M.indices, M.data,
                                       symmetric=symmetric, build_transpose=transpose)


def fixtures():
    rng = np.random.default_rng(7)
    mats = {"1138bus": load_mtx(mtx("1138bus")), "jpwh_991": load_mtx(mtx("jpwh_991")),
            "GD97_b": load_mtx(mtx("GD97_b"))}
    ip, ix, dv = kr.poisson2d_csr(123)
    mats["poisson2d_123"] = CsrRef((123 * 123, 123 * 123), ip, ix, dv)
    ip, ix, dv = kr.convdiff3d_csr(21)
    mats["convdiff3d_21"] = CsrRef((21 ** 3, 21 ** 3), ip, ix, dv)
    R = sp.random(3000, 2500, density=0.004, random_state=2, format="csr")
    R.sort_indices()
    mats["random_rect"] = CsrRef.from_scipy(R)
    # ragged: empty rows, one very long row (forces the thread-per-row fallback), unsorted columns
    n = 6000
    rows = [np.array([], dtype=np.int64)] * n
    rows = list(rows)
    rows[5] = rng.permutation(n)[:5000]
    rows[17] = np.array([3, 1, 2])
    rows[n - 1] = np.array([0, n - 1])
    for i in range(100, 400):
        rows[i] = rng.integers(0, n, size=rng.integers(0, 9))
    indptr = np.concatenate([[0], np.cumsum([len(r) for r in rows])]).astype(np.int32)
    indices = np.concatenate(rows).astype(np.int32)
    mats["ragged"] = CsrRef((n, n), indptr, indices, rng.standard_normal(len(indices)))
    mats["empty"] = CsrRef((50, 40), np.zeros(51, np.int32), np.zeros(0, np.int32), np.zeros(0))
    return mats


KERNELS = [("row", 1, 0, 0), ("stream", 2, 4096, 256), ("stream_small", 2, 512, 64),
           ("tma", 3, 2048, 256), ("tma_small", 3, 256, 128),
           ("rowpf", 6, 0, 0), ("rowpf2", 7, 0, 0)]     # software-pipelined row kernels (candidates)


# ============================================================== integer work
@pytest.mark.parametrize("name", ["1138bus", "jpwh_991", "GD97_b"])
def test_device_csr_equals_scipy_and_golden(ctx, name, golden):
    from pykrylov_b200.mmio import read_mtx
    shape, ip, ix, dv, sym = read_mtx(mtx(name))
    A = dev().DeviceCsr.from_arrays(ctx, shape, ip, ix, dv, symmetric=sym, build_transpose=True)
    p, i, d = A.download()
    g = golden["csr/" + name]
    assert zlib.crc32(p.tobytes()) == g["indptr_crc"] and zlib.crc32(i.tobytes()) == g["indices_crc"]
    assert zlib.crc32(d.tobytes()) == g["data_crc"]
    M = load_mtx(mtx(name))
    T = CsrRef.from_scipy(M.to_scipy().T.tocsr())
    tp, ti, td = A.download(transposed=True)
    assert np.array_equal(tp, T.indptr) and np.array_equal(ti, T.indices) and np.array_equal(td, T.data)
    assert np.array_equal(A.diagonal(), M.to_scipy().diagonal())


def test_device_transpose_of_unsymmetric_is_bit_exact(ctx):
    for name in ("jpwh_991", "random_rect", "ragged", "convdiff3d_21"):
        M = fixtures()[name]
        A = upload(ctx, M, transpose=True)
        tp, ti, td = A.download(transposed=True)
        S = M.to_scipy()
        # stable counting sort: inside a row of A^T the original row order is kept
        coo_r = np.repeat(np.arange(M.shape[0]), np.diff(M.indptr))
        order = np.argsort(M.indices, kind="stable")
        assert np.array_equal(ti, coo_r[order].astype(np.int32)) and np.array_equal(td, M.data[order])
        assert np.array_equal(tp, np.concatenate([[0], np.cumsum(np.bincount(M.indices, minlength=M.shape[1]))]))
        del S


def test_device_gallery_equals_oracle_generators(ctx):
    D = dev().DeviceCsr
    for g in (1, 2, 7, 64):
        p, i, d = D.poisson2d(ctx, g).download()
        rp, ri, rd = kr.poisson2d_csr(g)
        assert np.array_equal(p, rp) and np.array_equal(i, ri) and np.array_equal(d, rd)
    g = 40                                               # a row shard keeps global column ids
    p, i, d = D.poisson2d(ctx, g, 500, 1100).download()
    rp, ri, rd = kr.poisson2d_csr(g, 500, 1100)
    assert np.array_equal(p, rp) and np.array_equal(i, ri) and np.array_equal(d, rd)
    for m in (1, 3, 9):
        p, i, d = D.convdiff3d(ctx, m, 0.5).download()
        rp, ri, rd = kr.convdiff3d_csr(m, 0.5)
        assert np.array_equal(p, rp) and np.array_equal(i, ri) and np.array_equal(d, rd)
    p, i, d = D.poisson1d(ctx, 11).download()
    assert list(p[:3]) == [0, 2, 5] and list(d[:5]) == [2, -1, -1, 2, -1] and len(d) == 31


# ===================================================================== SpMV
@pytest.mark.parametrize("kname,kind,tile,threads", KERNELS)
def test_spmv_bit_exact_all_fixtures(ctx, kname, kind, tile, threads):
    rng = np.random.default_rng(0)
    for name, M in fixtures().items():
        A = upload(ctx, M, transpose=True)
        A.set_kernel(kind, tile, threads)
        x = rng.standard_normal(M.shape[1])
        xt = rng.standard_normal(M.shape[0])
        assert np.array_equal(A.matvec(x), M.matvec(x)), (name, kname)
        assert np.array_equal(A.matvec(xt, trans=True), M.rmatvec(xt)), (name, kname, "T")


@pytest.mark.parametrize("kname,kind,tile,threads", KERNELS)
def test_spmv_fused_dots(ctx, kname, kind, tile, threads):
    rng = np.random.default_rng(1)
    for name in ("jpwh_991", "poisson2d_123", "ragged"):
        M = fixtures()[name]
        A = upload(ctx, M)
        A.set_kernel(kind, tile, threads)
        x = rng.standard_normal(M.shape[1])
        w1, w2 = rng.standard_normal(M.shape[0]), rng.standard_normal(M.shape[0])
        xv, yv = ctx.vector(x), ctx.vector(M.shape[0])
        A.spmv_dot(xv, yv, [ctx.vector(w1), None, ctx.vector(w2)], slot0=5)
        y = M.matvec(x)
        assert np.array_equal(yv.download(), y)
        got = ctx.scalars(5, 3)
        for g_, ref in zip(got, (np.dot(w1, y), np.dot(y, y), np.dot(w2, y))):
            assert abs(g_ - ref) <= 1e-13 * np.sqrt(np.dot(y, y) * max(np.dot(w1, w1), np.dot(y, y)))


@pytest.mark.parametrize("kname,kind,tile,threads", KERNELS)
def test_spmv_with_fused_y_side_update(ctx, kname, kind, tile, threads):
    """kry_spmv_axpby_dot: z = a (A x) + b w with the coefficient rules of the multi-AXPY (immediate,
    slot, negated, dividing) in the SpMV's epilogue.  z is bit-identical to the two-launch form
    (SpMV into a temporary, then the vector pass); the fused inner product agrees to rounding."""
    rng = np.random.default_rng(7)
    D = dev()
    for name, trans in (("jpwh_991", False), ("random_rect", False), ("random_rect", True), ("ragged", False)):
        M = fixtures()[name]
        A = upload(ctx, M, transpose=trans)
        A.set_kernel(kind, tile, threads)
        nin, nout = (M.shape[0], M.shape[1]) if trans else (M.shape[1], M.shape[0])
        x, w0, q = rng.standard_normal(nin), rng.standard_normal(nout), rng.standard_normal(nout)
        ctx.set_scalars(10, [0.37, -1.25])
        for op_kw in (dict(a=1.0, b_slot=10, b_neg=1), dict(a_slot=11, b=2.5), dict(a_slot=10, a_div=True, b_slot=11, b_div=True),
                      dict(a=-3.0, w=None)):
            xv, tv = ctx.vector(x), ctx.vector(nout)
            z1, z2, qv = ctx.vector(w0), ctx.vector(w0), ctx.vector(q)
            has_w = "w" not in op_kw
            kw = {k: v for k, v in op_kw.items() if k != "w"}
            # two launches
            A.spmv(xv, tv, trans=trans)
            D.multi_axpy_dot(ctx, [dict(z=z1, u=tv, w=z1 if has_w else None, **kw)], [(z1, z1)], slot0=20)
            D.multi_axpy_dot(ctx, [], [(qv, z1)], slot0=21)
            # one launch (z aliases w), z.z, then q.z
            D.spmv_axpby_dot(A, xv, dict(z=z2, w=z2 if has_w else None, **kw), dot=True, slot0=22, trans=trans)
            ref = z1.download()
            assert np.array_equal(z2.download(), ref), (name, trans, op_kw)
            z3 = ctx.vector(w0)
            D.spmv_axpby_dot(A, xv, dict(z=z3, w=z3 if has_w else None, **kw), dot_with=qv, slot0=23, trans=trans)
            assert np.array_equal(z3.download(), ref)
            z4 = ctx.vector(w0)
            D.spmv_axpby_dot(A, xv, dict(z=z4, w=z4 if has_w else None, **kw), trans=trans)       # no inner product
            assert np.array_equal(z4.download(), ref)
            zz, qz, zz_f, qz_f = ctx.scalars(20, 4)
            assert abs(zz_f - zz) <= 1e-13 * zz and abs(zz - np.dot(ref, ref)) <= 1e-13 * zz
            assert abs(qz_f - qz) <= 1e-13 * np.sqrt(zz * np.dot(q, q))
    M = fixtures()["jpwh_991"]
    A = upload(ctx, M)
    xv, zv = ctx.vector(M.shape[1]), ctx.vector(M.shape[0])
    with pytest.raises(ValueError):
        D.spmv_axpby_dot(A, xv, dict(z=zv, w=xv))                     # w aliases the gathered vector
    with pytest.raises(L().KrylovDeviceError):
        D.spmv_axpby_dot(A, xv, dict(z=xv))                           # z aliases x
    with pytest.raises(ValueError):
        D.spmv_axpby_dot(A, xv, dict(z=ctx.vector(M.shape[0] + 1)))


def test_spmv_shape_errors_are_valueerror(ctx):
    M = fixtures()["random_rect"]
    A = upload(ctx, M)
    with pytest.raises(ValueError):
        A.matvec(np.ones(M.shape[1] + 1))
    with pytest.raises(ValueError):
        A.spmv(ctx.vector(M.shape[1]), ctx.vector(M.shape[0] + 2))
    with pytest.raises(L().KrylovDeviceError):
        A.matvec(np.ones(M.shape[0]), trans=True)           # transpose was not built
    with pytest.raises(L().KrylovDeviceError):
        dev().DeviceCsr.from_arrays(ctx, (2, 2), [0, 1, 3], [0, 1], [1.0, 1.0])   # rowptr[-1] != nnz


# =============================================================== multi-AXPY
def test_multi_axpy_dot_matches_numpy(ctx):
    rng = np.random.default_rng(3)
    n = 100003
    a, b, c = (rng.standard_normal(n) for _ in range(3))
    va, vb, vc = ctx.vector(a), ctx.vector(b), ctx.vector(c)
    ctx.set_scalars(10, [0.37, -1.25])
    ops = [dict(z=va, u=va, w=vb, a=1.0, b=0.37),                     # a += 0.37 b      (immediate)
           dict(z=vc, u=vc, w=va, a_slot=11, b=1.0, b_slot=10, b_neg=1),  # c = -1.25 c - 0.37 a (slots)
           dict(z=vb, u=vb, a=2.0)]                                   # b *= 2
    dev().multi_axpy_dot(ctx, ops, [(va, vc), (vb, vb)], slot0=20)
    a2 = a + 0.37 * b
    c2 = -1.25 * c + (-0.37) * a2
    b2 = 2.0 * b
    assert np.array_equal(va.download(), a2) and np.array_equal(vc.download(), c2)
    assert np.array_equal(vb.download(), b2)
    got = ctx.scalars(20, 2)
    assert srel(got[0], np.dot(a2, c2)) <= 1e-12 and srel(got[1], np.dot(b2, b2)) <= 1e-13
    with pytest.raises(ValueError):
        dev().multi_axpy_dot(ctx, [dict(z=va, u=ctx.vector(5), a=1.0)])


def test_reduction_is_deterministic(ctx):
    M = fixtures()["poisson2d_123"]
    A = upload(ctx, M)
    x = np.random.default_rng(5).standard_normal(M.shape[1])
    xv, yv = ctx.vector(x), ctx.vector(M.shape[0])
    seen = set()
    for _ in range(5):
        A.spmv_dot(xv, yv, [xv, None], slot0=0)
        seen.add(tuple(ctx.scalars(0, 2)))
    assert len(seen) == 1


# ==================================================== single-step solver parity
@pytest.fixture(params=[(0, 0), (1, 0), (2, 0), (2, 1)], ids=["cg3launch", "cgfuse1", "cgfuse2", "cg1cta"])
def cg_form(ctx, request):
    """Every CG test runs under the three multi-CTA launch plans (KRY_OPT_CG_FUSE: the 3-launch
    form and the two fused 2-launch forms) and with the one-CTA loop for problems that fit one SM
    (KRY_OPT_CG_ONE_CTA); all of them must be indistinguishable from outside."""
    d_fuse, d_one = ctx.get_option(L().KRY_OPT_CG_FUSE), ctx.get_option(L().KRY_OPT_CG_ONE_CTA)
    ctx.set_option(L().KRY_OPT_CG_FUSE, request.param[0])
    ctx.set_option(L().KRY_OPT_CG_ONE_CTA, request.param[1])
    yield request.param
    ctx.set_option(L().KRY_OPT_CG_FUSE, d_fuse)
    ctx.set_option(L().KRY_OPT_CG_ONE_CTA, d_one)


def cg_case(name):
    M = fixtures()[name]
    n = M.shape[0]
    rng = np.random.default_rng(21)
    return M, M.matvec(np.ones(n)), rng.standard_normal(n)


@pytest.mark.parametrize("name,precon", [("1138bus", 0), ("poisson2d_123", 0), ("1138bus", 1), ("1138bus", 2)])
def test_cg_single_step_from_identical_state(ctx, name, precon, cg_form):
    M, rhs, guess = cg_case(name)
    n = M.shape[0]
    d = np.abs(M.to_scipy().diagonal())
    pvec = None if precon == 0 else (1.0 / d if precon == 1 else d)
    pfun = None if precon == 0 else ((lambda r: pvec * r) if precon == 1 else (lambda r: r / pvec))
    A = upload(ctx, M, symmetric=True)
    st = kr.cg_start(M, rhs, guess=guess.copy(), precon=pfun, matvec_max=10 ** 6)
    S = dev().DeviceSolver(ctx, "cg", A)
    S.set_precon_diag(pvec, precon)
    S.setup(rhs, guess=guess, matvec_max=10 ** 6)
    d0 = S.status()
    assert srel(d0.resid_norm0, st.residNorm0) <= RTOL_STEP and d0.n_matvec == 1
    assert rel(S.get_vector("r"), st.r) <= RTOL_STEP and np.array_equal(S.get_vector("p"), -S.get_vector("r"))
    for k in range(12):
        for _ in range(3 if k else 0):
            kr.cg_step(M, st)
        # transplant the oracle state, then advance both by exactly one iteration
        S.set_vector("x", st.x); S.set_vector("r", st.r); S.set_vector("p", st.p)
        S.set_scalar("ry", float(st.ry))
        kr.cg_step(M, st)
        S.iterate(1)
        assert np.array_equal(S.get_vector("Ap"), st.Ap)           # SpMV: bit-exact
        within(srel(S.get_scalar("pAp"), st.pAp), RTOL_STEP, "cg:scalar pAp")
        within(srel(S.get_scalar("alpha"), st.alpha), RTOL_STEP, "cg:scalar alpha")
        within(srel(S.get_scalar("beta"), st.beta), RTOL_STEP, "cg:scalar beta")
        within(srel(S.get_scalar("ry"), st.ry), RTOL_STEP, "cg:scalar ry")
        within(srel(S.status().resid_norm, st.residNorm), RTOL_STEP, "cg:scalar residNorm")
        for v in ("x", "r", "p"):
            within(rel(S.get_vector(v), st[v]), RTOL_STEP, "cg:vector " + v)


def nonsym_case():
    M = fixtures()["jpwh_991"]
    n = M.shape[0]
    rng = np.random.default_rng(22)
    return M, M.matvec(rng.standard_normal(n)), rng.standard_normal(n)


@pytest.mark.parametrize("precon", [0, 2])
def test_bicgstab_single_step_from_identical_state(ctx, precon):
    M, rhs, guess = nonsym_case()
    d = np.maximum(np.abs(M.to_scipy().diagonal()), 1.0)
    pfun = None if precon == 0 else (lambda r: r / d)
    A = upload(ctx, M)
    st = kr.bicgstab_start(M, rhs, guess=guess.copy(), precon=pfun, matvec_max=10 ** 6)
    S = dev().DeviceSolver(ctx, "bicgstab", A)
    S.set_precon_diag(None if precon == 0 else d, precon)
    S.setup(rhs, guess=guess, matvec_max=10 ** 6)
    within(srel(S.status().resid_norm0, st.residNorm0), RTOL_STEP, "bicgstab:scalar residNorm")
    within(rel(S.get_vector("r0"), st.r0), RTOL_STEP, "bicgstab:vector r0")
    for k in range(10):
        for _ in range(2 if k else 0):
            kr.bicgstab_step(M, st)
        # device boundary state already holds the direction of the coming trip
        beta = st.rho_next / st.rho * st.alpha / st.omega
        p_new = st.p * beta
        p_new -= beta * st.omega * st.v
        p_new += st.r
        S.set_vector("x", st.x); S.set_vector("r", st.r); S.set_vector("r0", st.r0)
        S.set_vector("p", p_new); S.set_vector("v", st.v)
        if precon:
            S.set_vector("q", p_new / d)
        S.set_scalar("rho", float(st.rho_next))
        r_scale = np.max(np.abs(st.r))
        kr.bicgstab_step(M, st)
        S.iterate(1)
        assert np.array_equal(S.get_vector("v"), st.v)             # SpMV of identical input: bit-exact
        if st.finished:
            break
        alpha_err = srel(S.get_scalar("alpha"), st.alpha)
        # s = r - alpha v cancels: its relative error is alpha's (a dot product, accurate to
        # its own conditioning) amplified by |r|/|s|; t = A M s inherits it
        s_vec = st.s if precon else st.s / st.omega                # (z *= omega aliases s, :135)
        amp = max(1.0, r_scale / np.max(np.abs(s_vec)))
        tol = (RTOL_STEP + 4 * alpha_err) * amp
        within(alpha_err, RTOL_STEP, "bicgstab:scalar alpha")
        MARGINS["bicgstab:cancellation factor |r|/|s|"] = {"max_err": max(amp, MARGINS.get(
            "bicgstab:cancellation factor |r|/|s|", {"max_err": 0.0})["max_err"]), "bar": float("inf")}
        within(rel(S.get_vector("t"), st.t), tol, "bicgstab:vector t (bar scaled by |r|/|s|)")
        within(srel(S.get_scalar("omega"), st.omega), tol, "bicgstab:scalar omega (bar scaled by |r|/|s|)")
        within(srel(S.get_scalar("rho"), st.rho_next), 10 * tol, "bicgstab:scalar rho (-omega r0.t: one more cancellation)")
        within(srel(S.status().resid_norm, st.residNorm), tol, "bicgstab:scalar residNorm (bar scaled by |r|/|s|)")
        within(rel(S.get_vector("x"), st.x), RTOL_STEP, "bicgstab:vector x")  # x is not formed by cancellation
        assert rel(S.get_vector("r"), st.r) / max(1.0, np.max(np.abs(s_vec)) / np.max(np.abs(st.r))) <= tol, k


@pytest.mark.parametrize("precon", [0, 2])
def test_cgs_single_step_from_identical_state(ctx, precon):
    M, rhs, guess = nonsym_case()
    d = np.maximum(np.abs(M.to_scipy().diagonal()), 1.0)
    pfun = None if precon == 0 else (lambda r: r / d)
    A = upload(ctx, M)
    st = kr.cgs_start(M, rhs, guess=guess.copy(), precon=pfun, matvec_max=10 ** 6)
    S = dev().DeviceSolver(ctx, "cgs", A)
    S.set_precon_diag(None if precon == 0 else d, precon)
    S.setup(rhs, guess=guess, matvec_max=10 ** 6)
    assert S.status().n_matvec == 0                                # cgs.py:59-60: not counted
    for k in range(10):
        for _ in range(2 if k else 0):
            kr.cgs_step(M, st)
        for v in ("x", "r", "r0", "u", "p"):
            S.set_vector(v, st[v])
        if precon:
            S.set_vector("y", st.p / d)
        S.set_scalar("rho", float(st.rho))
        kr.cgs_step(M, st)
        S.iterate(1)
        within(srel(S.get_scalar("alpha"), st.alpha), RTOL_STEP, "cgs:scalar alpha")
        if st.finished:                      # beta / u / p are not formed on the last trip
            within(rel(S.get_vector("x"), st.x), RTOL_STEP, "early exit:vector x")
            assert S.status().done
            break
        within(srel(S.get_scalar("beta"), st.beta), RTOL_STEP, "cgs:scalar beta")
        within(srel(S.get_scalar("rho"), st.rho), RTOL_STEP, "cgs:scalar rho")
        within(srel(S.status().resid_norm, st.residNorm), RTOL_STEP, "cgs:scalar residNorm")
        for v in ("x", "r", "u", "p"):
            within(rel(S.get_vector(v), st[v]), RTOL_STEP, "cgs:vector " + v)


@pytest.mark.parametrize("precon", [0, 2])
def test_tfqmr_single_step_from_identical_state(ctx, precon):
    M, rhs, guess = nonsym_case()
    d = np.maximum(np.abs(M.to_scipy().diagonal()), 1.0)
    pfun = None if precon == 0 else (lambda r: r / d)
    A = upload(ctx, M)
    st = kr.tfqmr_start(M, rhs, guess=guess.copy(), precon=pfun, matvec_max=10 ** 6)
    S = dev().DeviceSolver(ctx, "tfqmr", A)
    S.set_precon_diag(None if precon == 0 else d, precon)
    S.setup(rhs, guess=guess, matvec_max=10 ** 6)
    assert S.status().n_matvec == 1 and np.array_equal(S.get_vector("u"), S.get_vector("v"))
    for k in range(10):
        for _ in range(2 if k else 0):
            kr.tfqmr_step(M, st)
        for v in ("x", "r0", "y", "w", "d", "u", "v"):
            S.set_vector(v, st[v])
        if precon:
            S.set_vector("z", st.z)
        S.set_scalar("rho", float(st.rho)); S.set_scalar("theta", float(st.theta))
        S.set_scalar("eta", float(st.eta)); S.set_scalar("resid", float(st.residNorm))
        S.set_scalar("k", float(st.k + 1))
        S.set_scalar("alpha", float(st.rho / np.dot(st.r0, st.v)))
        kr.tfqmr_step(M, st)
        S.iterate(1)
        if st.finished:
            within(rel(S.get_vector("x"), st.x), RTOL_STEP, "early exit:vector x")
            assert S.status().done
            break
        within(srel(S.get_scalar("theta"), st.theta), RTOL_STEP, "tfqmr:scalar theta")
        within(srel(S.get_scalar("eta"), st.eta), RTOL_STEP, "tfqmr:scalar eta")
        within(srel(S.get_scalar("rho"), st.rho), RTOL_STEP, "tfqmr:scalar rho")
        within(srel(S.status().resid_norm, st.residNorm), RTOL_STEP, "tfqmr:scalar residNorm")
        for v in ("x", "y", "w", "d", "u", "v"):
            within(rel(S.get_vector(v), st[v]), RTOL_STEP, "tfqmr:vector " + v)


def minres_case():
    M = fixtures()["jpwh_991"]
    S0 = M.to_scipy()
    Sm = CsrRef.from_scipy((S0 + S0.T) * 0.5)
    return Sm, Sm.matvec(np.ones(Sm.shape[0]))


@pytest.fixture(params=[0, 1, 2], ids=["minres3launch", "minres2launch", "minrespersistent"])
def minres_plan(ctx, request):
    """MINRES tests run under every launch plan: 3 launches, 2 launches (KRY_OPT_MINRES_FUSE) and the
    cooperative persistent kernel (KRY_OPT_MINRES_PERSISTENT; CUDA or the emulation's SIMT mode only)."""
    saved = (ctx.get_option(L().KRY_OPT_MINRES_FUSE), ctx.get_option(L().KRY_OPT_MINRES_PERSISTENT))
    ctx.set_option(L().KRY_OPT_MINRES_FUSE, 1 if request.param == 1 else 0)
    ctx.set_option(L().KRY_OPT_MINRES_PERSISTENT, 1 if request.param == 2 else 0)
    if request.param == 2 and not ctx.get_option(L().KRY_OPT_MINRES_PERSISTENT):
        ctx.set_option(L().KRY_OPT_MINRES_FUSE, saved[0])
        pytest.skip("the persistent kernel needs CUDA (or the SIMT mode of the emulation)")
    yield request.param
    ctx.set_option(L().KRY_OPT_MINRES_FUSE, saved[0])
    ctx.set_option(L().KRY_OPT_MINRES_PERSISTENT, saved[1])


@pytest.mark.parametrize("shift", [0.0, 0.5])
def test_minres_single_step_from_identical_state(ctx, shift, minres_plan):
    M, rhs = minres_case()
    A = upload(ctx, M, symmetric=True)
    st = kr.minres_start(M, rhs, shift=shift)
    S = dev().DeviceSolver(ctx, "minres", A)
    S.setup(rhs, matvec_max=10 ** 6, shift=shift, rtol=1e-12, etol=1e-6, window=5)
    within(srel(S.status().resid_norm0, st.beta1), RTOL_STEP, "minres:scalar residNorm")
    names = dict(mbeta="beta", oldb="oldb", dbar="dbar", epsln="epsln", phibar="phibar", cs="cs", sn="sn",
                 tnorm2="tnorm2", ynorm2="ynorm2", rhs1="rhs1", rhs2="rhs2", gmax="gmax", gmin="gmin",
                 beta1="beta1", xnrg2="xNrgNorm2")
    for k in range(8):
        for _ in range(3 if k else 0):
            kr.minres_step(M, st)
        assert st.istop == 0
        S.set_scalar("n_iter", st.itn); S.set_scalar("n_matvec", st.itn)
        S.set_vector("r2", st.r2); S.set_vector("r1", st.r1 if st.itn else st.r2)
        S.set_vector("w", st.w); S.set_vector("w2", st.w2); S.set_vector("x", st.x)
        for dname, oname in names.items():
            S.set_scalar(dname, float(st[oname]))
        for j in range(5):
            S.set_scalar("derr%d" % j, float(st.dErr[j]))
        kr.minres_step(M, st)
        S.iterate(1)
        ds = S.status()
        assert ds.n_iter == st.itn and ds.istop == st.istop
        within(srel(S.get_scalar("alfa"), st.alfa), RTOL_STEP, "minres:scalar alfa")
        within(srel(S.get_scalar("mbeta"), st.beta), RTOL_STEP, "minres:scalar mbeta")
        for dname in ("phibar", "cs", "sn", "tnorm2", "ynorm2", "rhs1", "gmax", "gmin", "dbar", "epsln"):
            within(srel(S.get_scalar(dname), float(st[names.get(dname, dname)])), RTOL_STEP, "minres:scalar " + dname)
        within(srel(ds.resid_norm, st.rnorm), RTOL_STEP, "minres:scalar rnorm")
        within(srel(ds.aux[0], st.Anorm), RTOL_STEP, "minres:scalar Anorm")
        within(srel(ds.aux[3], st.Arnorm), RTOL_STEP, "minres:scalar Arnorm")
        for v in ("x", "r2", "r1", "w"):
            within(rel(S.get_vector(v), st[v]), RTOL_STEP, "minres:vector " + v)


# ================================================= trajectories and known answers
def test_cg_trajectory_poisson2d(ctx, golden, cg_form):
    g = 100
    ip, ix, dv = kr.poisson2d_csr(g)
    M = CsrRef((g * g, g * g), ip, ix, dv)
    rhs = M.matvec(np.ones(g * g))
    ref = kr.cg_solve(M, rhs)
    A = dev().DeviceCsr.poisson2d(ctx, g)
    S = dev().DeviceSolver(ctx, "cg", A)
    S.setup(rhs, matvec_max=2 * g * g)
    st = S.run(7)
    hist = S.drain_history(st)[:, 0]
    rec = golden["cg/poisson2d_csr/100"]
    assert st.n_matvec == ref.nMatvec == rec["nMatvec"] == 160
    within(rel(hist[:20], ref.residHistory[:20]), RTOL_STEP, "cg_trajectory_poisson2d:hist_20_ref_residHistory_20_")
    assert np.max(np.abs(hist - np.array(ref.residHistory)) / np.array(ref.residHistory)) <= 1e-10
    assert srel(st.resid_norm0, rec["residNorm0"]) <= 1e-14
    # final residual: true residuals of both solutions agree to 1e-8 (relative to |b|)
    x = S.solution()
    nb = np.linalg.norm(rhs)
    assert abs(np.linalg.norm(rhs - M.matvec(x)) - np.linalg.norm(rhs - M.matvec(ref.x))) / nb <= RTOL_FINAL
    assert rel(x, ref.x) <= 1e-10


def test_public_api_known_answers(ctx, golden):
    """The reference's own CG tests (pykrylov/cg/tests/test_diagdom.py) through the
    public classes, on device operators."""
    from pykrylov_b200.cg import CG
    from pykrylov_b200.gallery import poisson1d_operator, poisson2d_operator
    from pykrylov_b200.tools import machine_epsilon
    from math import sin, pi
    for n in (10, 20, 100, 1000, 5000):
        A = poisson1d_operator(n, context=ctx)
        e = np.ones(n)
        rhs = A * e
        cg = CG(A, matvec_max=2 * n, outputStream=None)           # ctor kwargs ignored like the reference
        cg.solve(rhs)
        cond = (4.0 * sin((n - 1) * pi / 2.0 / n) ** 2) / (4.0 * sin(pi / 2.0 / n) ** 2)
        assert cg.nMatvec == n // 2 and cg.converged
        assert np.allclose(e, cg.bestSolution, rtol=cond * machine_epsilon())
        assert A.nMatvec == 1 + cg.nMatvec
    for g, nmv in ((10, 15), (20, 33), (100, 160), (500, 756)):
        A = poisson2d_operator(g, context=ctx)
        e = np.ones(g * g)
        cg = CG(A)
        cg.solve(A * e)
        assert cg.nMatvec == nmv and cg.converged
        if g <= 100:
            rec = golden["cg/poisson2d_csr/%d" % g]
            assert srel(cg.residNorm0, rec["residNorm0"]) <= 1e-14
            assert len(cg.residHistory) == nmv + 1
            within(rel(cg.residHistory[:25], rec["residHistory"]), RTOL_STEP, "public_api_known_answers:cg_residHistory_25_rec_residHistory_")
        assert np.linalg.norm(e - cg.bestSolution) / g <= 1e-5


def test_bmark_through_pysparse_shim(ctx, golden):
    """examples/bmark.py:34-54, unchanged apart from print()."""
    from pykrylov.linop import PysparseLinearOperator
    from pykrylov.cgs import CGS
    from pykrylov.tfqmr import TFQMR
    from pykrylov.bicgstab import BiCGSTAB
    from pysparse import spmatrix
    from pysparse.sparse.pysparseMatrix import PysparseMatrix as spm
    AA = spmatrix.ll_mat_from_mtx(mtx("jpwh_991"))
    A = spm(matrix=AA)
    op = PysparseLinearOperator(A)
    n = A.shape[0]
    e = np.ones(n)
    rhs = A * e
    assert np.array_equal(rhs, load_mtx(mtx("jpwh_991")).matvec(e))
    for KSolver in [CGS, TFQMR, BiCGSTAB]:
        ks = KSolver(op, reltol=1.0e-8)
        ks.solve(rhs, guess=1 + np.arange(n, dtype=op.dtype), matvec_max=2 * n)
        err = np.linalg.norm(ks.bestSolution - e) / np.sqrt(n)
        rec = golden["%s/jpwh_991/reltol1e-08" % ks.acronym]
        assert ks.converged and abs(ks.nMatvec - rec["nMatvec"]) <= 4      # doc table: 82/84/84
        assert srel(ks.residNorm0, rec["residNorm0"]) <= 1e-13
        assert ks.residNorm <= 1e-8 * ks.residNorm0 and err <= 1e-5
        assert ks.residHistory == []


def test_short_runs_reproduce_doc_tables(ctx, golden):
    """doc/source/{cgs,bicgstab}.rst:50-52 (reltol 1e-5): identical counts and digits."""
    from pykrylov_b200.linop import csr_operator
    from pykrylov_b200.cgs import CGS
    from pykrylov_b200.bicgstab import BiCGSTAB
    M = load_mtx(mtx("jpwh_991"))
    n = M.shape[0]
    op = csr_operator(M.shape, M.indptr, M.indices, M.data, context=ctx)
    e = np.ones(n)
    for K, nmv, res, err in ((CGS, 64, "4.72e-03", "1.47e-04"), (BiCGSTAB, 57, "5.18e-02", "3.35e-03")):
        ks = K(op, reltol=1.0e-5)
        ks.solve(M.matvec(e), guess=1 + np.arange(n, dtype=float), matvec_max=2 * n)
        assert ks.nMatvec == nmv and "%8.2e" % ks.residNorm == res
        assert "%8.2e" % (np.linalg.norm(ks.bestSolution - e) / np.sqrt(n)) == err


def test_minres_public_api(ctx, golden, capsys, minres_plan):
    from pykrylov_b200.linop import csr_operator
    from pykrylov_b200.minres import Minres
    M, rhs = minres_case()
    op = csr_operator(M.shape, M.indptr, M.indices, M.data, symmetric=True, context=ctx)
    mr = Minres(op)
    mr.solve(rhs, show=False)
    g = golden["MINRES/sym_jpwh_991"]
    assert mr.istop == 10 and abs(mr.itn - g["itn"]) <= 2 and mr.converged and mr.nMatvec == mr.itn
    assert srel(mr.residNorm0, g["residNorm0"]) <= 1e-14
    assert rel(mr.residHistory[:15], g["residHistory"][:15]) <= 1e-9
    assert srel(mr.Acond, g["Acond"]) <= 1e-6 and np.linalg.norm(mr.x - 1) / np.sqrt(len(rhs)) <= 1e-5
    assert len(mr.dir_errors_window) == mr.itn - 5
    # nonsymmetric operator: the default symmetry check stops at 0 iterations, istop 7
    N = load_mtx(mtx("jpwh_991"))
    nop = csr_operator(N.shape, N.indptr, N.indices, N.data, context=ctx)
    mr = Minres(nop)
    mr.solve(N.matvec(np.ones(N.shape[0])), show=False)
    g = golden["MINRES/jpwh_991_nonsym"]
    assert (mr.istop, mr.itn) == (g["istop"], g["itn"]) and not mr.converged
    mr = Minres(op)
    mr.solve(rhs)                                                   # show=True prints the banner
    out = capsys.readouterr().out
    assert "Enter minres." in out and "istop   =   10" in out


# ================================================================ edge cases
def test_cg_edge_cases(ctx, cg_form):
    from pykrylov_b200.linop import csr_operator
    from pykrylov_b200.cg import CG
    # zero right-hand side: converged at entry, x = 0, no product
    ip, ix, dv = kr.poisson2d_csr(5)
    op = csr_operator((25, 25), ip, ix, dv, symmetric=True, context=ctx)
    cg = CG(op)
    cg.solve(np.zeros(25))
    assert cg.nMatvec == 0 and cg.converged and cg.residNorm == 0.0 and np.array_equal(cg.x, np.zeros(25))
    assert cg.residHistory == [0.0]
    # 1 x 1 system
    one = csr_operator((1, 1), [0, 1], [0], [4.0], symmetric=True, context=ctx)
    cg = CG(one)
    cg.solve(np.array([2.0]))
    assert cg.nMatvec == 1 and np.array_equal(cg.x, [0.5])
    # matvec_max stops the loop exactly like cg.py:113
    cg = CG(op)
    cg.solve(op * np.ones(25), matvec_max=3)
    assert cg.nMatvec == 3 and not cg.converged and len(cg.residHistory) == 4
    # indefinite operator: curvature test (cg.py:119-124)
    ind = csr_operator((2, 2), [0, 1, 2], [0, 1], [1.0, -1.0], symmetric=True, context=ctx)
    cg = CG(ind)
    cg.solve(np.array([1.0, 2.0]))
    ref = kr.cg_solve(CsrRef((2, 2), [0, 1, 2], [0, 1], [1.0, -1.0]), np.array([1.0, 2.0]))
    assert not cg.definite and not ref.definite and cg.nMatvec == ref.nMatvec
    assert np.array_equal(cg.infiniteDescent, ref.infiniteDescent) and np.array_equal(cg.x, ref.x)
    cg = CG(ind)
    cg.solve(np.array([1.0, 2.0]), check_curvature=False, matvec_max=4)
    ref = kr.cg_solve(CsrRef((2, 2), [0, 1, 2], [0, 1], [1.0, -1.0]), np.array([1.0, 2.0]),
                      check_curvature=False, matvec_max=4)
    assert cg.nMatvec == ref.nMatvec and np.allclose(cg.x, ref.x, rtol=1e-12, equal_nan=True)
    # check_symmetric on a nonsymmetric operator: logs and returns None without solving
    N = load_mtx(mtx("jpwh_991"))
    nop = csr_operator(N.shape, N.indptr, N.indices, N.data, context=ctx)
    cg = CG(nop)
    assert cg.solve(np.ones(991), check_symmetric=True) is None and cg.bestSolution is None
    # wrong-size right-hand side
    with pytest.raises(ValueError):
        CG(op).solve(np.ones(24))
    # store_iterates / store_resids keep per-iteration copies
    cg = CG(op)
    cg.solve(op * np.ones(25), store_iterates=True, store_resids=True)
    assert len(cg.iterates) == cg.nMatvec + 1 and len(cg.resids) == cg.nMatvec + 1


def test_cg_jacobi_precon_reproduces_reference_quirk(ctx, golden, cg_form):
    """cg.py:104,150-151 build p from r, not from the preconditioned residual; with a
    Jacobi preconditioner on 1138bus the reference therefore stalls until matvec_max."""
    from pykrylov_b200.linop import csr_operator, DiagonalOperator
    from pykrylov_b200.cg import CG
    M = load_mtx(mtx("1138bus"))
    op = csr_operator(M.shape, M.indptr, M.indices, M.data, symmetric=True, context=ctx)
    cg = CG(op, precon=DiagonalOperator(1.0 / M.to_scipy().diagonal()))
    cg.solve(M.matvec(np.ones(M.shape[0])))
    rec = golden["cg/1138bus/jacobi"]
    assert cg.nMatvec == rec["nMatvec"] and cg.converged == rec["converged"]
    assert srel(cg.residNorm0, rec["residNorm0"]) <= 1e-13
    assert rel(cg.residHistory[:10], rec["residHistory"][:10]) <= 1e-10


@pytest.mark.parametrize("name,precon,guess", [("poisson2d_123", 0, False), ("1138bus", 0, True),
                                               ("1138bus", 2, True), ("GD97_b", 1, False)])
def test_cg_fused_forms_are_bit_identical_to_the_3_launch_form(ctx, name, precon, guess):
    """The fused forms only move work between launches: after any number of iterations --
    read in the middle of a run or at the end, through CUDA-graph replays or single
    launches -- x, r, p, Ap, every scalar and the residual history are the same bits."""
    M = fixtures()[name]
    n = M.shape[0]
    rng = np.random.default_rng(31)
    rhs = M.matvec(np.ones(n))
    x0 = rng.standard_normal(n) if guess else None
    d = np.abs(M.to_scipy().diagonal()) + 1.0
    A = upload(ctx, M, symmetric=True)
    default = ctx.get_option(L().KRY_OPT_CG_FUSE)
    one_cta = ctx.get_option(L().KRY_OPT_CG_ONE_CTA)
    ctx.set_option(L().KRY_OPT_CG_ONE_CTA, 0)           # this test is about the multi-CTA plans
    out = {}
    try:
        for form in (0, 1, 2):
            ctx.set_option(L().KRY_OPT_CG_FUSE, form)
            S = dev().DeviceSolver(ctx, "cg", A)
            S.set_precon_diag(d if precon else None, precon)
            S.setup(rhs, guess=x0, abstol=0.0, reltol=0.0, matvec_max=10 ** 6, check_curvature=False)
            snaps = []
            for chunk in (1, 2, 40, 3, 31):          # 40 and 31 go through the graph replay path
                S.iterate(chunk)
                st = S.status()
                snaps.append((st.n_iter, st.n_matvec, st.resid_norm, tuple(st.aux[:4]),
                              S.get_vector("x"), S.get_vector("r"), S.get_vector("p"), S.get_vector("Ap")))
            snaps.append(S.drain_history())
            out[form] = snaps
            S._release()
    finally:
        ctx.set_option(L().KRY_OPT_CG_FUSE, default)
        ctx.set_option(L().KRY_OPT_CG_ONE_CTA, one_cta)
    for form in (1, 2):
        for a, b in zip(out[0][:-1], out[form][:-1]):
            assert a[:4] == b[:4], (form, a[:4], b[:4])
            for u, v in zip(a[4:], b[4:]):
                assert np.array_equal(u, v, equal_nan=True), form
        assert np.array_equal(out[0][-1], out[form][-1], equal_nan=True)


def test_bicgstab_breakdown_runs_to_matvec_max_like_reference(ctx):
    """rhs = A*ones with a zero guess is an exact breakdown on jpwh_991 (rho' = 0):
    the reference keeps iterating on NaNs until matvec_max; so must the device."""
    from pykrylov_b200.linop import csr_operator
    from pykrylov_b200.bicgstab import BiCGSTAB
    M = load_mtx(mtx("jpwh_991"))
    n = M.shape[0]
    op = csr_operator(M.shape, M.indptr, M.indices, M.data, context=ctx)
    ks = BiCGSTAB(op, reltol=1e-8)
    ks.solve(M.matvec(np.ones(n)), matvec_max=2 * n)
    with np.errstate(all="ignore"):
        ref = kr.bicgstab_solve(M, M.matvec(np.ones(n)), reltol=1e-8, matvec_max=2 * n)
    if np.isnan(ks.residNorm):
        assert ks.nMatvec == ref.nMatvec == 2 * n and not ks.converged
    else:                       # a different summation order may dodge the exact zero
        assert ks.converged


def test_closure_operator_goes_through_host_bridge(ctx):
    """The reference's own test operator: LinearOperator(n, n, lambda x: Poisson1dMatvec(x))."""
    from pykrylov_b200.linop import LinearOperator
    from pykrylov_b200.gallery import Poisson1dMatvec, Poisson2dMatvec
    from pykrylov_b200.cg import CG
    n = 100
    A = LinearOperator(n, n, lambda x: Poisson1dMatvec(x), symmetric=True)
    e = np.ones(n)
    cg = CG(A, context=ctx)
    cg.solve(A * e)
    assert cg.nMatvec == 50 and np.allclose(cg.bestSolution, e, rtol=1e-10)
    ref = kr.cg_solve(kr.poisson1d_matvec, kr.poisson1d_matvec(e))
    assert rel(cg.residHistory[:30], ref.residHistory[:30]) <= 1e-9
    A2 = LinearOperator(400, 400, lambda x: Poisson2dMatvec(x), symmetric=True)
    cg = CG(A2, context=ctx)
    cg.solve(A2 * np.ones(400))
    assert cg.nMatvec == 33 and cg.converged


# ===================================== full-size, size-independent properties
@pytest.fixture(scope="module")
def big(ctx):
    g = 3162                                       # BASELINE.json config 2: N = 9 998 244
    A = dev().DeviceCsr.poisson2d(ctx, g)
    return g, A


def test_fullsize_row_sums_are_exact(ctx, big):
    g, A = big
    n = g * g
    assert A.shape == (n, n) and A.nnz == 5 * n - 4 * g
    ones, y = dev().DeviceVector(ctx, n).fill(1.0), dev().DeviceVector(ctx, n)
    for kind, tile, thr in ((1, 0, 0), (2, 4096, 256), (3, 2048, 256)):
        A.set_kernel(kind, tile, thr)
        A.spmv_dot(ones, y, [ones, None], slot0=0)
        s = ctx.scalars(0, 2)
        # A*1 is 0 in the interior, 1 on edges, 2 in corners: sums are exact integers
        assert s[0] == 4.0 * g and s[1] == 4.0 * g + 8.0
    A.set_kernel(0, 0, 0)


def test_fullsize_kernels_agree_bitwise_and_operator_is_symmetric(ctx, big):
    g, A = big
    n = g * g
    rng = np.random.default_rng(9)
    x, w = ctx.vector(rng.standard_normal(n)), ctx.vector(rng.standard_normal(n))
    y1, y2, z = (dev().DeviceVector(ctx, n) for _ in range(3))
    A.set_kernel(1, 0, 0)
    A.spmv_dot(x, y1, [w], slot0=0)
    w_Ax = ctx.scalars(0, 1)[0]
    for kind, tile, thr in ((2, 4096, 256), (3, 2048, 512)):
        A.set_kernel(kind, tile, thr)
        A.spmv(x, y2)
        dev().multi_axpy_dot(ctx, [dict(z=z, u=y1, w=y2, a=1.0, b=-1.0)], [(z, z)], slot0=1)
        assert ctx.scalars(1, 1)[0] == 0.0                         # bit-identical vectors
    A.set_kernel(0, 0, 0)
    A.spmv_dot(w, y2, [x], slot0=2)
    x_Aw = ctx.scalars(2, 1)[0]
    within(srel(w_Ax, x_Aw), RTOL_STEP, "fullsize_kernels_agree_bitwise_and_operator_is_symmetric:w_Ax_x_Aw")  # symmetry, reduction-order noise only
    # linearity: A(2x - 3w) == 2Ax - 3Aw up to rounding of the combination
    dev().multi_axpy_dot(ctx, [dict(z=z, u=x, w=w, a=2.0, b=-3.0)])
    y3 = dev().DeviceVector(ctx, n)
    A.spmv(z, y3)
    dev().multi_axpy_dot(ctx, [dict(z=z, u=y1, w=y2, a=2.0, b=-3.0), dict(z=z, u=z, w=y3, a=1.0, b=-1.0)],
                         [(z, z), (y3, y3)], slot0=3)
    s = ctx.scalars(3, 2)
    assert np.sqrt(s[0] / s[1]) <= 1e-14


def test_fullsize_cg_residual_recurrence_is_consistent(ctx, big, cg_form):
    g, A = big
    n = g * g
    ones = dev().DeviceVector(ctx, n).fill(1.0)
    rhs = dev().DeviceVector(ctx, n)
    A.spmv(ones, rhs)
    S = dev().DeviceSolver(ctx, "cg", A)
    runs = []
    for _ in range(2):
        S.setup_dev(rhs, abstol=0.0, reltol=0.0, matvec_max=60)
        st = S.run(16)
        runs.append((st.resid_norm, st.n_matvec, st.n_iter))
    assert runs[0] == runs[1] and runs[0][1] == 60                 # deterministic, exactly K iterations
    # recurrence residual r = A x - b versus the explicitly recomputed one
    x = ctx.vector(S.solution())
    Ax, r = dev().DeviceVector(ctx, n), ctx.vector(S.get_vector("r"))
    A.spmv(x, Ax)
    z = dev().DeviceVector(ctx, n)
    dev().multi_axpy_dot(ctx, [dict(z=z, u=Ax, w=rhs, a=1.0, b=-1.0), dict(z=z, u=z, w=r, a=1.0, b=-1.0)],
                         [(z, z), (r, r), (rhs, rhs)], slot0=0)
    s = ctx.scalars(0, 3)
    assert np.sqrt(s[0] / s[2]) <= RTOL_FINAL                      # |(Ax-b) - r| / |b|
    assert srel(np.sqrt(s[1]), runs[0][0]) <= 1e-12


# ============================================================ lifetime / host staging
def test_handles_can_be_destroyed_in_any_order():
    """Interpreter shutdown finalises objects in arbitrary order: destroying the context
    before its vectors / operators / solvers must neither crash nor leak, and a call through
    a handle whose context is gone fails with a status code."""
    from pykrylov_b200.device import Context, DeviceCsr, DeviceSolver
    c = Context(0)
    A = DeviceCsr.poisson2d(c, 8)
    v = c.vector(np.ones(64))
    S = DeviceSolver(c, "cg", A)
    S.setup(np.ones(64))
    S.iterate(3)
    handles = (S._h, A._h, v._h)
    lib = L().lib
    lib.kry_ctx_destroy(c._h)                       # context first, children still alive
    c._h = L().handle()
    y = np.zeros(64)
    assert lib.kry_vec_download(v._h, y.ctypes.data, 64) == L().KRY_ERR_STATE
    assert lib.kry_solver_iterate(S._h, 1) == L().KRY_ERR_STATE
    assert "destroyed" in L().last_error()
    with pytest.raises(L().KrylovDeviceError):
        A.spmv(v, v)
    # children afterwards, in the "wrong" order; the last one frees the context struct
    lib.kry_vec_destroy(handles[2]); v._h = L().handle()
    lib.kry_csr_destroy(handles[1]); A._h = L().handle()
    lib.kry_solver_destroy(handles[0]); S._h = L().handle()
    # and the library is still healthy
    c2 = Context(0)
    ip, ix, dv = kr.poisson2d_csr(4)
    assert np.array_equal(DeviceCsr.poisson2d(c2, 4).matvec(np.ones(16)), CsrRef((16, 16), ip, ix, dv).matvec(np.ones(16)))
    c2.close()


def test_large_results_come_back_in_pooled_pinned_memory(ctx):
    from pykrylov_b200.device import result_pool
    n = 1 << 18                                     # 2 MiB: above the pool's threshold
    v = ctx.vector(np.arange(n, dtype=np.float64))
    a = v.download()
    assert type(a.base).__name__ == "_PinnedBlock" and np.array_equal(a, np.arange(n))
    hits = result_pool.hits
    del a
    b = v.download()                                # the block is recycled, not re-pinned
    assert result_pool.hits == hits + 1 and np.array_equal(b, np.arange(n))
    small = ctx.vector(np.ones(10)).download()
    assert small.base is None


def test_cg_one_cta_loop_matches_oracle_and_counts_one_launch_per_call(ctx, golden):
    """KRY_OPT_CG_ONE_CTA: 1138bus (90 kB) iterates inside one CTA -- one launch per
    kry_solver_iterate call, same history as the oracle within the per-step bar."""
    M = fixtures()["1138bus"]
    n = M.shape[0]
    rhs = M.matvec(np.ones(n))
    A = upload(ctx, M, symmetric=True)
    saved = ctx.get_option(L().KRY_OPT_CG_ONE_CTA)
    ctx.set_option(L().KRY_OPT_CG_ONE_CTA, 1)
    try:
        S = dev().DeviceSolver(ctx, "cg", A)
        S.setup(rhs, matvec_max=2 * n)
        l0 = ctx.launch_count()
        S.iterate(40)
        assert ctx.launch_count() - l0 == 1
        st = S.status()
        ref = kr.cg_solve(M, rhs, matvec_max=40)
        hist = S.drain_history(st)[:, 0]
        # cond(1138bus) ~ 1e7: a different summation order alone moves the history by 1e-15 at
        # trip 20, 1e-6 at trip 30 and 1e-2 at trip 40 (measured with the oracle, np.dot against
        # math.fsum), so the trajectory is compared over the first 20 trips
        assert st.n_matvec == 40 and rel(hist[:21], ref.residHistory[:21]) <= 1e-9
        st = S.run(200)
        full = kr.cg_solve(M, rhs)
        assert abs(st.n_matvec - full.nMatvec) <= 0.01 * full.nMatvec       # cond ~ 1e7: SURVEY.md section 6
        x = S.solution()
        assert abs(np.linalg.norm(rhs - M.matvec(x)) - np.linalg.norm(rhs - M.matvec(full.x))) <= RTOL_FINAL * np.linalg.norm(rhs)
    finally:
        ctx.set_option(L().KRY_OPT_CG_ONE_CTA, saved)


def test_pipelined_host_drive_gives_the_same_results(ctx):
    """_engine.drive(overlap=True) keeps one chunk in flight (status snapshots through
    kry_solver_status_enqueue/_wait, history on the copy stream); results, histories and
    counters must not depend on it or on the check interval."""
    import pykrylov_b200._engine as eng
    from pykrylov_b200.linop import csr_operator
    from pykrylov_b200.cg import CG
    from pykrylov_b200.bicgstab import BiCGSTAB
    from pykrylov_b200.minres import Minres
    M = load_mtx(mtx("1138bus"))
    op = csr_operator(M.shape, M.indptr, M.indices, M.data, symmetric=True, context=ctx)
    rhs = M.matvec(np.ones(M.shape[0]))
    N = load_mtx(mtx("jpwh_991"))
    nop = csr_operator(N.shape, N.indptr, N.indices, N.data, context=ctx)
    nrhs = N.matvec(np.random.default_rng(3).standard_normal(N.shape[0]))
    real = eng.drive
    out = {}
    try:
        for overlap in (False, True):
            eng.drive = (lambda S, k, cb=None, overlap=True, _o=overlap: real(S, k, cb, overlap=_o and overlap))
            for interval in (7, 64):
                cg = CG(op, check_interval=interval)
                cg.solve(rhs)
                bi = BiCGSTAB(nop, reltol=1e-8, check_interval=interval)
                bi.solve(nrhs, matvec_max=2 * N.shape[0])
                mr = Minres(op, check_interval=interval)
                mr.solve(rhs, show=False, check=False, itnlim=300)
                out[(overlap, interval)] = (cg.nMatvec, tuple(cg.residHistory), cg.bestSolution.copy(),
                                            bi.nMatvec, tuple(bi.residHistory), bi.bestSolution.copy(),
                                            mr.itn, mr.istop, tuple(mr.residHistory), mr.x.copy())
    finally:
        eng.drive = real
    base = out[(False, 7)]
    for key, val in out.items():
        for a, b in zip(base, val):
            assert (np.array_equal(a, b) if isinstance(a, np.ndarray) else a == b), key


def test_minres_2_launch_plan_is_bit_identical_to_the_3_launch_plan(ctx):
    """KRY_OPT_MINRES_FUSE only moves the w / x update of a trip into the next trip's second
    launch: x, w, w2, r1, r2, every scalar and the history are the same bits, read mid-run
    (which settles what is owed) or at the end, through graph replays or single launches."""
    S0 = fixtures()["jpwh_991"].to_scipy()
    M = CsrRef.from_scipy(((S0 + S0.T) * 0.5).tocsr())
    n = M.shape[0]
    rhs = M.matvec(np.ones(n))
    A = upload(ctx, M, symmetric=True)
    default = ctx.get_option(L().KRY_OPT_MINRES_FUSE)
    out = {}
    try:
        for plan in (0, 1):
            ctx.set_option(L().KRY_OPT_MINRES_FUSE, plan)
            S = dev().DeviceSolver(ctx, "minres", A)
            S.setup(rhs, matvec_max=10 ** 6, rtol=0.0, etol=0.0, window=5)
            snaps = []
            for chunk in (1, 2, 40, 3, 31):
                S.iterate(chunk)
                st = S.status()
                snaps.append((st.n_iter, st.resid_norm, tuple(st.aux[:13])) +
                             tuple(S.get_vector(v) for v in ("x", "w", "w2", "r1", "r2")))
            snaps.append(S.drain_history())
            out[plan] = snaps
            S._release()
    finally:
        ctx.set_option(L().KRY_OPT_MINRES_FUSE, default)
    for a, b in zip(out[0][:-1], out[1][:-1]):
        assert a[:3] == b[:3], (a[:3], b[:3])
        for u, v in zip(a[3:], b[3:]):
            assert np.array_equal(u, v, equal_nan=True)
    assert np.array_equal(out[0][-1], out[1][-1], equal_nan=True)


@pytest.mark.parametrize("pmode", [1, 2])
def test_minres_with_diagonal_preconditioner_is_device_resident(ctx, pmode):
    """Minres.solve(precon=<diagonal>) iterates on the device (no host bridge) and follows the
    reference loop: same stop reason and iteration count, history and solution to rounding."""
    import pykrylov_b200._engine as eng
    from pykrylov_b200.linop import DiagonalOperator, csr_operator
    from pykrylov_b200.minres import Minres
    S0 = fixtures()["jpwh_991"].to_scipy()
    M = CsrRef.from_scipy(((S0 + S0.T) * 0.5).tocsr())
    n = M.shape[0]
    rhs = M.matvec(np.ones(n))
    d = 0.5 + np.random.default_rng(12).random(n)
    ref = kr.minres_solve(M, rhs, precon=(lambda r: d * r) if pmode == 1 else (lambda r: r / d))
    op = csr_operator(M.shape, M.indptr, M.indices, M.data, symmetric=True, context=ctx)

    class DiagonalPrec(object):                     # examples/bmark.py:14-22
        def __init__(self, diag):
            self.diag = diag

        def __call__(self, y):
            return y / self.diag

        __mul__ = __call__

    precon = DiagonalOperator(d) if pmode == 1 else DiagonalPrec(d)
    bridged = []
    real = eng.HostBridge

    class Spy(real):
        def __init__(self, *a, **k):
            bridged.append(1)
            real.__init__(self, *a, **k)

    eng.HostBridge = Spy
    try:
        mr = Minres(op, context=ctx)
        mr.solve(rhs, precon=precon, show=False, check=False)
    finally:
        eng.HostBridge = real
    assert not bridged
    assert mr.istop == ref.istop and abs(mr.itn - ref.itn) <= 1
    k = min(len(ref.residHistory), len(mr.residHistory), 20)
    assert rel(mr.residHistory[:k], ref.residHistory[:k]) <= 1e-9
    assert np.linalg.norm(mr.x - ref.x) <= 1e-6 * np.linalg.norm(ref.x)


# ------------------------------------------------------------------ BASELINE configs 3 and 4 at full size
@pytest.mark.gpu
def test_config3_minres_kron_jpwh_fullsize_follows_the_oracle(ctx):
    """BASELINE config 3 at its full size (N = 999 919, nnz = 6 404 123): MINRES on
    kron(I_1009, sym(jpwh_991)) -- SpMV bit-exact on a random vector, and under every launch plan
    the oracle's exit (istop 10 at 74 iterations, the reference's own figures), its
    residual history and truncated direct-error estimates, and its solution."""
    from pykrylov_b200.gallery import kron_sym_jpwh
    shape, ip, ix, dv = kron_sym_jpwh(mtx("jpwh_991"), 1009)
    n = shape[0]
    assert (n, len(dv)) == (999919, 6404123)
    M = CsrRef(shape, ip, ix, dv)
    A = dev().DeviceCsr.from_arrays(ctx, shape, ip, ix, dv, symmetric=True)
    v = np.random.default_rng(5).standard_normal(n)
    assert np.array_equal(A.matvec(v), M.matvec(v))
    rhs = M.matvec(np.ones(n))
    ref = kr.minres_solve(M, rhs)
    assert (ref.istop, ref.itn) == (10, 74)
    rh = np.array(ref.residHistory)
    de = np.array(ref.dir_errors_window)
    bnorm = np.linalg.norm(rhs)

    def oracle_residual_after(itn):
        st = kr.minres_start(M, rhs)
        while st.itn < itn:
            kr.minres_step(M, st)
        return np.linalg.norm(rhs - M.matvec(st.x))

    # How far the trajectory can be followed.  The Lanczos recurrence loses orthogonality on this
    # operator: re-running the ORACLE with nothing changed but the summation order of its inner
    # products (reversed, or in chunks of 991) moves its own history by 1e-10 at trip 40, 1e-6 at
    # trip 45, 1e-3 at trip 49 and 1-8 % at the exit, where the truncated direct-error estimate sits
    # within 1.5 % of etol one trip before the exit.  So: per-entry agreement
    # over the first 30 trips, the oracle's exit reason, and its exit trip +- 1.
    saved = [ctx.get_option(o) for o in (L().KRY_OPT_MINRES_PERSISTENT, L().KRY_OPT_MINRES_FUSE)]
    runs = {}
    try:
        for name, pers, fuse in (("3-launch", 0, 0), ("2-launch", 0, 1), ("persistent", 1, 0)):
            ctx.set_option(L().KRY_OPT_MINRES_PERSISTENT, pers)
            ctx.set_option(L().KRY_OPT_MINRES_FUSE, fuse)
            S = dev().DeviceSolver(ctx, "minres", A)
            S.setup(rhs, abstol=0.0, reltol=0.0, matvec_max=5 * n, rtol=1e-12, etol=1e-6, window=5)
            st = S.run(16)
            h = S.drain_history(st)
            itn = int(st.n_iter)
            assert int(st.istop) == ref.istop and abs(itn - ref.itn) <= 1, (name, st.istop, itn)
            assert len(h) == itn and rel(h[:30, 0], rh[:30]) <= 1e-10, (name, rel(h[:30, 0], rh[:30]))
            d = h[:, 1][~np.isnan(h[:, 1])]
            assert len(d) == itn - 5 and rel(d[:25], de[:25]) <= 1e-10, name
            x = S.solution()
            res = np.linalg.norm(rhs - M.matvec(x))
            assert abs(res - oracle_residual_after(itn)) <= RTOL_FINAL * bnorm, name     # final residual bar
            assert np.max(np.abs(x - 1.0)) <= 1e-4                                       # exact solution: ones
            runs[name] = (h.copy(), x.copy())
            S._release()
    finally:
        ctx.set_option(L().KRY_OPT_MINRES_PERSISTENT, saved[0])
        ctx.set_option(L().KRY_OPT_MINRES_FUSE, saved[1])
    # the 2-launch plan only moves the w / x update between launches
    assert rel(runs["2-launch"][0][:30, 0], runs["3-launch"][0][:30, 0]) <= 1e-12


@pytest.mark.gpu
def test_config4_bicgstab_convdiff_215_fullsize_follows_the_oracle(ctx):
    """BASELINE config 4 at its full size (7-point convection-diffusion, grid 215^3, N = 9 938 375,
    nnz = 69 291 275): the device-generated operator equals the oracle's CSR, A x and A^T x are
    bit-exact on a random vector, and the first Bi-CGSTAB iterations follow the oracle."""
    m = 215
    n = m ** 3
    ip, ix, dv = kr.convdiff3d_csr(m)
    assert (n, len(dv)) == (9938375, 69291275)
    M = CsrRef((n, n), ip, ix, dv)
    A = dev().DeviceCsr.convdiff3d(ctx, m, 0.5, build_transpose=True)
    dip, dix, ddv = A.download()
    assert np.array_equal(dip, ip) and np.array_equal(dix, ix) and np.array_equal(ddv, dv)
    del dip, dix, ddv
    v = np.random.default_rng(6).standard_normal(n)
    assert np.array_equal(A.matvec(v), M.matvec(v))
    assert np.array_equal(A.matvec(v, trans=True), M.rmatvec(v))
    rhs = M.matvec(np.ones(n))
    ref = kr.bicgstab_solve(M, rhs, reltol=1e-8, matvec_max=10)
    rh = np.array(ref.residHistory)
    S = dev().DeviceSolver(ctx, "bicgstab", A)
    S.setup(rhs, abstol=1e-8, reltol=1e-8, matvec_max=10)
    st = S.run(8)
    h = S.drain_history(st)[:, 0]
    assert st.n_matvec == ref.nMatvec == 10
    assert len(h) == len(rh) and np.max(np.abs(h - rh) / rh) <= 1e-10, np.max(np.abs(h - rh) / rh)
    x = S.solution()
    assert np.max(np.abs(x - ref.x)) <= 1e-10 * np.max(np.abs(ref.x))
    S._release()
    A._release()


# ------------------------------------------------------------------ device-side assembly and operator algebra
def _reference_coo_matvec(vals, rows, cols, nargout, x, symmetric):
    """linop/linop.py:647-664 restated (the reference's Python loop)."""
    y = np.zeros(nargout)
    for k in range(len(vals)):
        y[rows[k]] += vals[k] * x[cols[k]]
        if symmetric and rows[k] != cols[k]:
            y[cols[k]] += vals[k] * x[rows[k]]
    return y


@pytest.mark.parametrize("symmetric", [False, True])
def test_coord_operator_is_assembled_on_the_device_in_the_reference_order(ctx, symmetric):
    """CoordLinearOperator (linop.py:638-685): COO -> CSR runs in HBM (stable sort by row, symmetric
    expansion, range check); inside a row the entries keep the reference loop's accumulation order,
    so A x and A^T y equal that loop bit for bit -- duplicates, empty rows and all."""
    from pykrylov_b200.linop import CoordLinearOperator, CsrLinearOperator
    rng = np.random.default_rng(12)
    m, n = (60, 60) if symmetric else (50, 70)
    nnz = 900
    rows = rng.integers(0, m, nnz)
    cols = rng.integers(0, n, nnz)
    if symmetric:
        rows, cols = np.maximum(rows, cols), np.minimum(rows, cols)       # one triangle (with diagonal entries)
    rows[rows == 7] = 8                                                  # an empty row
    vals = rng.standard_normal(nnz)
    op = CoordLinearOperator(vals, rows, cols, nargin=n, nargout=m, symmetric=symmetric, context=ctx)
    assert isinstance(op, CsrLinearOperator) and op.shape == (m, n) and op.symmetric == symmetric
    ip, ix, dv = op.device_csr.download()
    # the CSR itself: per row, entries in arrival order (entry k before its mirror image of a later k)
    want = [[] for _ in range(m)]
    for k in range(nnz):
        want[rows[k]].append((cols[k], vals[k]))
        if symmetric and rows[k] != cols[k]:
            want[cols[k]].append((rows[k], vals[k]))
    assert np.array_equal(ip, np.concatenate([[0], np.cumsum([len(w) for w in want])]))
    assert np.array_equal(ix, np.array([c for w in want for c, _ in w], dtype=np.int32))
    assert np.array_equal(dv, np.array([v for w in want for _, v in w]))
    x = rng.standard_normal(n)
    assert np.array_equal(op * x, _reference_coo_matvec(vals, rows, cols, m, x, symmetric))
    y = rng.standard_normal(m)
    assert np.array_equal(op.T * y, _reference_coo_matvec(vals, cols, rows, n, y, symmetric))   # linop.py:666-681
    dense = np.zeros((m, n))
    for k in range(nnz):
        dense[rows[k], cols[k]] += vals[k]
        if symmetric and rows[k] != cols[k]:
            dense[cols[k], rows[k]] += vals[k]
    assert np.array_equal(op.to_array(), dense)
    with pytest.raises(IndexError):
        CoordLinearOperator(vals, rows + m, cols, nargin=n, nargout=m, symmetric=symmetric, context=ctx)
    # the reference's own test cases (linop/tests/test_linop.py:374-396)
    A = rng.random((5, 7))
    r, c = np.nonzero(A)
    C1 = CoordLinearOperator(A[r, c], r, c, 7, 5, symmetric=False, context=ctx)
    xx, yy = rng.random(7), rng.random(5)
    assert np.allclose(A @ xx, C1 * xx, rtol=1e-14) and np.allclose(A.T @ yy, C1.T * yy, rtol=1e-14)


def test_operator_algebra_on_device_operators_stays_in_hbm(ctx):
    """A + sigma*I, A - D, -A, A/2, 3*A, A + B, I - A, A*B of device operators are device operators
    (reference linop.py:307-345, 378-410 returns host closures): solvers keep their device-resident
    loop.  Shifts, diagonal terms and negation reproduce the reference's closure expression bit for
    bit; general scalings and sums to rounding (documented in csrc/assemble.cu)."""
    import pykrylov_b200._engine as eng
    from pykrylov_b200.linop import (CsrLinearOperator, DeviceChainOperator, DiagonalOperator, IdentityOperator,
                                     LinearOperator, csr_operator)
    from pykrylov_b200.cg import CG
    rng = np.random.default_rng(21)
    M = fixtures()["poisson2d_123"]
    n = M.shape[0]
    A = csr_operator(M.shape, M.indptr, M.indices, M.data, symmetric=True, context=ctx)
    R = sp.random(n, n, density=0.001, random_state=3, format="csr")
    R.sort_indices()
    Rm = CsrRef.from_scipy(R)
    B = csr_operator(R.shape, R.indptr, R.indices, R.data, context=ctx)
    x = rng.standard_normal(n)
    d = rng.standard_normal(n)
    D, I = DiagonalOperator(d), IdentityOperator(n)
    Ax, Bx = M.matvec(x), Rm.matvec(x)

    def dev(op):
        assert isinstance(op, CsrLinearOperator) and op.device_csr is not None, type(op)
        return op

    # exact cases: the reference's closures evaluate A(v) + sign*other(v) and alpha*A(v)
    assert np.array_equal(dev(A + 0.75 * I) * x, Ax + 0.75 * x)
    assert np.array_equal(dev(A - 0.75 * I) * x, Ax + (-1) * (0.75 * x))
    assert np.array_equal(dev(A + I) * x, Ax + x)
    assert np.array_equal(dev(A - D) * x, Ax + (-1) * (d * x))
    assert np.array_equal(dev(D + A) * x, d * x + Ax)
    assert np.array_equal(dev(I - A) * x, x + (-1) * Ax)
    assert np.array_equal(dev(-A) * x, -1 * Ax)
    assert np.array_equal(dev(A / 2) * x, 0.5 * Ax) and np.array_equal(dev(4 * A) * x, 4 * Ax)
    # rounding-level cases
    tol = 1e-14 * np.max(np.abs(Ax))
    assert np.max(np.abs(dev(3 * A) * x - 3 * Ax)) <= 3 * tol
    assert np.max(np.abs(dev(A * 0.3) * x - 0.3 * Ax)) <= tol
    assert np.max(np.abs(dev(A + B) * x - (Ax + Bx))) <= tol
    assert np.max(np.abs(dev(A - B) * x - (Ax - Bx))) <= tol
    S = dev(A + B)
    assert not S.symmetric and dev(A + A).symmetric
    assert np.max(np.abs(S.T * x - (M.rmatvec(x) + Rm.rmatvec(x)))) <= tol
    # products: SpMVs back to back in HBM
    P = A * B
    assert isinstance(P, DeviceChainOperator) and P.shape == (n, n)
    assert np.array_equal(P * x, M.matvec(Bx))
    assert np.array_equal(P.T * x, Rm.rmatvec(M.rmatvec(x)))
    assert np.array_equal((P * A) * x, M.matvec(Rm.matvec(Ax)))
    assert type(A * 0) .__name__ == "ZeroOperator"
    # anything the device cannot see keeps the reference's closure semantics
    H = LinearOperator(n, n, lambda v: 2 * v, symmetric=True)
    mixed = A + H
    assert not isinstance(mixed, CsrLinearOperator) and np.array_equal(mixed * x, Ax + 2 * x)
    with pytest.raises(Exception):
        A + csr_operator((n, n + 1), np.zeros(n + 1, np.int32), np.zeros(0, np.int32), np.zeros(0), context=ctx)
    # a shifted operator handed to a solver iterates on the device
    shifted = A + 0.5 * I
    rhs = shifted * np.ones(n)
    assert eng.resolve(shifted, None, n) is not None
    cg = CG(shifted)
    l0 = ctx.launch_count()
    cg.solve(rhs)
    ip, ix, dv = shifted.device_csr.download()
    ref = kr.cg_solve(CsrRef((n, n), ip, ix, dv), rhs)
    assert cg.nMatvec == ref.nMatvec and ctx.launch_count() - l0 >= 2 * cg.nMatvec
    assert rel(np.array(cg.residHistory), np.array(ref.residHistory)) <= 1e-9
    assert np.max(np.abs(cg.bestSolution - 1.0)) <= 1e-5
    # to_array on the device
    small = csr_operator((3, 4), np.array([0, 2, 2, 4]), np.array([0, 3, 1, 1]), np.array([1.0, 2.0, 3.0, 4.0]), context=ctx)
    assert np.array_equal(small.to_array(), np.array([[1.0, 0, 0, 2.0], [0, 0, 0, 0], [0, 7.0, 0, 0]]))
