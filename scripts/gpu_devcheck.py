#!/usr/bin/env python3
"""Development check run under gpurun: parity of every kernel / solver against the
oracle on small cases + an SpMV / CG timing sweep.  Writes gpurun_out/devcheck.json.
Not part of the test-suite (tests/ holds the real parity tests); this is the
fast "does the first path work on a B200" probe."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import krylov_ref as kr           # noqa: E402
from oracle.csr_ref import CsrRef, load_mtx   # noqa: E402
from pykrylov_b200 import _lib as L           # noqa: E402
from pykrylov_b200.device import Context, DeviceCsr, DeviceSolver, DeviceVector  # noqa: E402

OUT = {}
GOLD = os.path.join(ROOT, "tests", "golden")


def relerr(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    d = np.linalg.norm(a - b)
    n = np.linalg.norm(b)
    return float(d / n) if n > 0 else float(d)


def spmv_parity(ctx):
    res = {}
    rng = np.random.default_rng(0)
    mats = {"1138bus": load_mtx(os.path.join(GOLD, "1138bus.mtx")),
            "jpwh_991": load_mtx(os.path.join(GOLD, "jpwh_991.mtx"))}
    ip, ix, dv = kr.poisson2d_csr(300)
    mats["poisson2d_300"] = CsrRef((90000, 90000), ip, ix, dv)
    ip, ix, dv = kr.convdiff3d_csr(30)
    mats["convdiff_30"] = CsrRef((27000, 27000), ip, ix, dv)
    import scipy.sparse as sp
    R = sp.random(5000, 4000, density=0.002, random_state=1, format="csr")
    R.sort_indices()
    mats["random_rect"] = CsrRef.from_scipy(R)
    for name, M in mats.items():
        A = DeviceCsr.from_arrays(ctx, M.shape, M.indptr, M.indices, M.data, build_transpose=True)
        x = rng.standard_normal(M.shape[1])
        xt = rng.standard_normal(M.shape[0])
        yref, ytref = M.matvec(x), M.rmatvec(xt)
        for kind, kname in ((L.KRY_SPMV_ROW, "row"), (L.KRY_SPMV_STREAM, "stream"), (L.KRY_SPMV_TMA, "tma")):
            for tile, thr in ((4096, 256), (1024, 128)):
                A.set_kernel(kind, tile, thr)
                y = A.matvec(x)
                yt = A.matvec(xt, trans=True)
                # fused dots
                xv, yv = ctx.vector(x), ctx.vector(M.shape[0])
                wv = ctx.vector(rng.standard_normal(M.shape[0]))
                A.spmv_dot(xv, yv, [wv, None], slot0=3)
                sc = ctx.scalars(3, 2)
                w = wv.download()
                key = "%s/%s/%d" % (name, kname, tile)
                res[key] = dict(bit_exact=bool(np.array_equal(y, yref)),
                                bit_exact_T=bool(np.array_equal(yt, ytref)),
                                dot_rel=[abs(sc[0] - np.dot(w, yref)) / max(abs(np.dot(w, yref)), 1e-300),
                                         abs(sc[1] - np.dot(yref, yref)) / np.dot(yref, yref)])
        # device transpose == scipy's
        tp, ti, td = A.download(transposed=True)
        T = CsrRef.from_scipy(M.to_scipy().T.tocsr())
        res[name + "/transpose_equal"] = bool(np.array_equal(tp, T.indptr) and np.array_equal(ti, T.indices)
                                              and np.array_equal(td, T.data))
    # device gallery == oracle generators
    for g in (7, 64):
        A = DeviceCsr.poisson2d(ctx, g)
        p, i, d = A.download()
        rp, ri, rd = kr.poisson2d_csr(g)
        res["gallery_poisson2d_%d" % g] = bool(np.array_equal(p, rp) and np.array_equal(i, ri) and np.array_equal(d, rd))
    A = DeviceCsr.convdiff3d(ctx, 9)
    p, i, d = A.download()
    rp, ri, rd = kr.convdiff3d_csr(9)
    res["gallery_convdiff_9"] = bool(np.array_equal(p, rp) and np.array_equal(i, ri) and np.array_equal(d, rd))
    return res


def solver_parity(ctx):
    res = {}
    # CG on Poisson2D grid 100, rhs = A*ones
    g = 100
    ip, ix, dv = kr.poisson2d_csr(g)
    M = CsrRef((g * g, g * g), ip, ix, dv)
    rhs = M.matvec(np.ones(g * g))
    ref = kr.cg_solve(M, rhs)
    A = DeviceCsr.poisson2d(ctx, g)
    S = DeviceSolver(ctx, "cg", A)
    S.setup(rhs, matvec_max=2 * g * g)
    st = S.run(16)
    hist = S.drain_history(st)
    k = min(len(hist), len(ref.residHistory))
    res["cg_poisson2d_100"] = dict(nMatvec=[int(st.n_matvec), int(ref.nMatvec)],
                                   resid=[st.resid_norm, ref.residNorm],
                                   hist_rel_max=float(np.max(np.abs(hist[:k, 0] - np.array(ref.residHistory[:k])) /
                                                             np.array(ref.residHistory[:k]))),
                                   x_rel=relerr(S.solution(), ref.x))
    # CG with guess on 1138bus: first 30 iterations
    M = load_mtx(os.path.join(GOLD, "1138bus.mtx"))
    n = M.shape[0]
    rhs = M.matvec(np.ones(n))
    guess = 1.0 + np.arange(n, dtype=float)
    stt = kr.cg_start(M, rhs, guess=guess.copy())
    for _ in range(30):
        kr.cg_step(M, stt)
    A = DeviceCsr.from_arrays(ctx, M.shape, M.indptr, M.indices, M.data, symmetric=True)
    S = DeviceSolver(ctx, "cg", A)
    S.setup(rhs, guess=guess, matvec_max=2 * n)
    S.iterate(30)
    st = S.status()
    hist = S.drain_history(st)
    res["cg_1138bus_30"] = dict(hist_rel_max=float(np.max(np.abs(hist[:31, 0] - np.array(stt.residHistory[:31])) /
                                                          np.array(stt.residHistory[:31]))),
                                x_rel=relerr(S.solution(), stt.x), nMatvec=[int(st.n_matvec), stt.nMatvec])
    # nonsymmetric solvers on jpwh_991, bmark settings
    M = load_mtx(os.path.join(GOLD, "jpwh_991.mtx"))
    n = M.shape[0]
    rhs = M.matvec(np.ones(n))
    guess = 1.0 + np.arange(n, dtype=float)
    A = DeviceCsr.from_arrays(ctx, M.shape, M.indptr, M.indices, M.data)
    for name, solve in (("bicgstab", kr.bicgstab_solve), ("cgs", kr.cgs_solve), ("tfqmr", kr.tfqmr_solve)):
        for tag, gs in (("guess", guess), ("zero", None)):
            # (rhs = A*ones with a zero guess is an exact Bi-CGSTAB breakdown on this matrix)
            b = rhs if gs is not None else M.matvec(np.random.default_rng(5).standard_normal(n))
            with np.errstate(all="ignore"):
                ref = solve(M, b, guess=None if gs is None else gs.copy(), reltol=1e-8, matvec_max=2 * n)
            S = DeviceSolver(ctx, name, A)
            S.setup(b, guess=gs, reltol=1e-8, matvec_max=2 * n)
            st = S.run(8)
            res["%s_jpwh_%s" % (name, tag)] = dict(nMatvec=[int(st.n_matvec), int(ref.nMatvec)],
                                                   resid=[st.resid_norm, float(ref.residNorm)],
                                                   resid0=[st.resid_norm0, float(ref.residNorm0)],
                                                   converged=[int(st.converged), int(bool(ref.converged))],
                                                   x_rel=relerr(S.solution(), ref.x))
    # MINRES on sym(jpwh_991)
    Ssp = M.to_scipy()
    Ssym = CsrRef.from_scipy((Ssp + Ssp.T) * 0.5)
    rhs = Ssym.matvec(np.ones(n))
    ref = kr.minres_solve(Ssym, rhs)
    A = DeviceCsr.from_arrays(ctx, Ssym.shape, Ssym.indptr, Ssym.indices, Ssym.data, symmetric=True)
    S = DeviceSolver(ctx, "minres", A)
    S.setup(rhs, matvec_max=5 * n, rtol=1e-12, etol=1e-6, window=5, shift=0.0)
    st = S.run(8)
    hist = S.drain_history(st)
    k = min(len(hist), len(ref.residHistory))
    res["minres_symjpwh"] = dict(istop=[int(st.istop), int(ref.istop)], itn=[int(st.n_iter), int(ref.itn)],
                                 rnorm=[st.resid_norm, float(ref.rnorm)],
                                 hist_rel_max=float(np.max(np.abs(hist[:k, 0] - np.array(ref.residHistory[:k])) /
                                                           np.array(ref.residHistory[:k]))),
                                 Anorm=[st.aux[0], ref.Anorm], Acond=[st.aux[1], ref.Acond],
                                 x_rel=relerr(S.solution(), ref.x))
    return res


def timing(ctx):
    res = {}
    g = 3162
    n = g * g
    t0 = time.time()
    A = DeviceCsr.poisson2d(ctx, g)
    ctx.sync()
    res["gen_poisson2d_3162_s"] = time.time() - t0
    nnz = A.nnz
    spmv_bytes = 12 * nnz + 4 * (n + 1) + 16 * n
    x = DeviceVector(ctx, n).fill(1.0)
    y = DeviceVector(ctx, n)
    configs = [(L.KRY_SPMV_ROW, 0, 0, "row")]
    for tile in (2048, 4096, 8192):
        for thr in (128, 256, 512):
            configs.append((L.KRY_SPMV_STREAM, tile, thr, "stream"))
            configs.append((L.KRY_SPMV_TMA, tile, thr, "tma"))
    for kind, tile, thr, name in configs:
        A.set_kernel(kind, tile, thr)
        try:
            for _ in range(3):
                A.spmv_dot(x, y, [x], slot0=0)
            best = 1e9
            for _ in range(3):
                ctx.flush_l2()
                ctx.timer_start()
                for _ in range(10):
                    A.spmv_dot(x, y, [x], slot0=0)
                best = min(best, ctx.timer_stop() / 10)
            res["spmv_dot/%s/t%d/b%d" % (name, tile, thr)] = dict(ms=best, GBs=spmv_bytes / best / 1e6,
                                                                   pAp=float(ctx.scalars(0, 1)[0]))
        except Exception as e:       # keep sweeping
            res["spmv_dot/%s/t%d/b%d" % (name, tile, thr)] = dict(error=str(e))
    # CG iterations/s with the default kernel
    A.set_kernel(L.KRY_SPMV_AUTO, 0, 0)
    rhs = DeviceVector(ctx, n)
    A.spmv(x, rhs)
    S = DeviceSolver(ctx, "cg", A)
    for kind, name in ((L.KRY_SPMV_STREAM, "stream"), (L.KRY_SPMV_TMA, "tma"), (L.KRY_SPMV_ROW, "row")):
        A.set_kernel(kind, 4096, 256)
        S.setup_dev(rhs, abstol=0.0, reltol=0.0, matvec_max=10 ** 9)
        S.iterate(20)
        ctx.sync()
        ctx.timer_start()
        S.iterate(200)
        ms = ctx.timer_stop()
        st = S.status()
        res["cg_iters_per_s/" + name] = dict(ms_per_iter=ms / 200, iters_per_s=200 / (ms / 1e3),
                                             GBs=(spmv_bytes + 72 * n) / (ms / 200) / 1e6,
                                             resid=st.resid_norm, n_iter=int(st.n_iter))
    # plain vector bandwidth yardstick: device memcpy
    ctx.timer_start()
    for _ in range(10):
        y.copy_from(x)
    ms = ctx.timer_stop() / 10
    res["memcpy_d2d_GBs"] = 16 * n / ms / 1e6
    return res


def main():
    ctx = Context(0)
    OUT["props"] = ctx.props()
    for name, fn in (("spmv_parity", spmv_parity), ("solver_parity", solver_parity), ("timing", timing)):
        if len(sys.argv) > 1 and name not in sys.argv[1:]:
            continue
        t0 = time.time()
        try:
            OUT[name] = fn(ctx)
        except Exception as e:
            import traceback
            OUT[name] = dict(error=repr(e), tb=traceback.format_exc())
        OUT[name + "_seconds"] = time.time() - t0
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "devcheck.json"), "w") as fh:
        json.dump(OUT, fh, indent=1, default=float)
    print(json.dumps(OUT, indent=1, default=float))


if __name__ == "__main__":
    main()
