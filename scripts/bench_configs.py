#!/usr/bin/env python3
"""Secondary measurements: BASELINE.json configs[0], [2], [3] (bench.py covers [1] and [4])
plus the stand-alone A^T x SpMV.  One JSON line per config; the CPU reference arm is the
transliterated reference (oracle/_ref) when present, else the oracle port.

  configs[0]  CG on 1138bus.mtx (latency-bound: N = 1138)
  configs[2]  MINRES on kron(I_1009, (B+B^T)/2), B = jpwh_991  (N = 999 919)
  configs[3]  Bi-CGSTAB on the 7-pt convection-diffusion grid 215^3 (N = 9 938 375) + A^T x
"""
import io
import json
import os
import sys
import time
from contextlib import redirect_stdout

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle import krylov_ref as kr                          # noqa: E402
from pykrylov_b200.device import Context, DeviceCsr, DeviceSolver, DeviceVector   # noqa: E402
from pykrylov_b200.mmio import read_mtx                      # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def ref_classes():
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if os.path.isdir(os.path.join(ref_dir, "refpykrylov")):
        sys.path.insert(0, ref_dir)
        import refpykrylov.linop as lo
        from refpykrylov.cg import CG
        from refpykrylov.bicgstab import BiCGSTAB
        from refpykrylov.minres import Minres
        return lo, dict(cg=CG, bicgstab=BiCGSTAB, minres=Minres)
    return None, None


def time_device(ctx, S, setup, steps, warmup):
    setup()
    S.iterate(warmup)
    ctx.sync()
    ctx.timer_start()
    S.iterate(steps)
    ms = ctx.timer_stop()
    st = S.status()
    assert st.n_iter == warmup + steps, (st.n_iter, warmup, steps, st.done)
    return ms / steps, st


def emit(**kw):
    print(json.dumps(kw), flush=True)


def note(msg):
    print("[bench_configs] " + msg, file=sys.stderr, flush=True)


def config0(ctx):
    import scipy.sparse as sp
    shape, ip, ix, dv, sym = read_mtx(os.path.join(GOLD, "1138bus.mtx"))
    n = shape[0]
    A = DeviceCsr.from_arrays(ctx, shape, ip, ix, dv, symmetric=True)
    M = sp.csr_matrix((dv, ix, ip), shape=shape)
    rhs = M @ np.ones(n)
    S = DeviceSolver(ctx, "cg", A)
    # throughput: fixed number of iterations
    ms, _ = time_device(ctx, S, lambda: S.setup(rhs, abstol=0.0, reltol=0.0, matvec_max=10 ** 9), 2000, 50)
    # the reference's own run: defaults, to convergence, through the public API (e2e)
    from pykrylov_b200.linop import CsrLinearOperator
    from pykrylov_b200.cg import CG
    op = CsrLinearOperator(A)
    CG(op).solve(rhs)
    t0 = time.perf_counter()
    cg = CG(op, check_interval=64)
    cg.solve(rhs)
    e2e_s = time.perf_counter() - t0
    lo, K = ref_classes()
    t0 = time.perf_counter()
    if K:
        ref = K["cg"](lo.LinearOperator(n, n, lambda v: M @ v, symmetric=True))
        ref.solve(rhs)
        nmv_ref, kind = ref.nMatvec, "reference"
    else:
        ref = kr.cg_solve(lambda v: M @ v, rhs)
        nmv_ref, kind = ref.nMatvec, "port"
    cpu_s = time.perf_counter() - t0
    emit(config="configs[0]: CG on 1138bus.mtx fp64 (N=1138, nnz=4054), defaults", metric="cg_iters_per_s",
         value=1e3 / ms, unit="iters/s", ms_per_step=ms,
         e2e={"solve_s": e2e_s, "nMatvec": cg.nMatvec, "iters_per_s": cg.nMatvec / e2e_s,
              "residNorm": cg.residNorm, "converged": bool(cg.converged)},
         cpu_baseline={"kind": kind, "solve_s": cpu_s, "nMatvec": int(nmv_ref), "iters_per_s": nmv_ref / cpu_s},
         note="latency-bound: 153 kB per iteration; time is launch + reduction latency, not bandwidth")


def config2(ctx):
    import scipy.sparse as sp
    shape, ip, ix, dv, _ = read_mtx(os.path.join(GOLD, "jpwh_991.mtx"))
    B = sp.csr_matrix((dv, ix, ip), shape=shape)
    S1 = ((B + B.T) * 0.5).tocsr()
    S1.sort_indices()
    k = 1009
    m = shape[0]
    n = k * m
    nnz1 = S1.nnz
    indptr = (np.arange(k, dtype=np.int64)[:, None] * nnz1 + S1.indptr[None, :-1]).reshape(-1)
    indptr = np.concatenate([indptr, [k * nnz1]]).astype(np.int32)
    indices = (np.arange(k, dtype=np.int64)[:, None] * m + S1.indices[None, :]).reshape(-1).astype(np.int32)
    data = np.tile(S1.data, k)
    A = DeviceCsr.from_arrays(ctx, (n, n), indptr, indices, data, symmetric=True)
    M = sp.csr_matrix((data, indices, indptr), shape=(n, n))
    rhs = M @ np.ones(n)
    S = DeviceSolver(ctx, "minres", A)
    ms, _ = time_device(ctx, S, lambda: S.setup(rhs, matvec_max=10 ** 9, rtol=0.0, etol=0.0, window=5), 60, 20)
    it_bytes = 12 * A.nnz + 4 * (n + 1) + 16 * n + 96 * n
    # to convergence through the public API
    from pykrylov_b200.linop import CsrLinearOperator
    from pykrylov_b200.minres import Minres
    op = CsrLinearOperator(A)
    t0 = time.perf_counter()
    Minres(op).solve(rhs, show=False, check=False)      # first solve on this operator: solver slab, pinned result block
    e2e_first_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    mr = Minres(op)
    mr.solve(rhs, show=False, check=False)
    e2e_s = time.perf_counter() - t0
    lo, K = ref_classes()
    t0 = time.perf_counter()
    if K:
        ref = K["minres"](lo.LinearOperator(n, n, lambda v: M @ v, symmetric=True))
        with redirect_stdout(io.StringIO()):
            ref.solve(rhs, show=False, check=False)
        itn_ref, istop_ref, kind = ref.itn, ref.istop, "reference"
    else:
        ref = kr.minres_solve(lambda v: M @ v, rhs)
        itn_ref, istop_ref, kind = ref.itn, ref.istop, "port"
    cpu_s = time.perf_counter() - t0
    emit(config="configs[2]: MINRES fp64 on kron(I_1009, sym(jpwh_991)) (N=%d, nnz=%d)" % (n, A.nnz),
         metric="minres_iters_per_s", value=1e3 / ms, unit="iters/s", ms_per_step=ms,
         achieved_GBs=it_bytes / ms / 1e6, algorithmic_bytes_per_step=it_bytes,
         e2e={"solve_s": e2e_s, "first_solve_s": e2e_first_s, "itn": mr.itn, "istop": mr.istop,
              "iters_per_s": mr.itn / e2e_s, "rnorm": mr.rnorm},
         cpu_baseline={"kind": kind, "solve_s": cpu_s, "itn": int(itn_ref), "istop": int(istop_ref),
                       "iters_per_s": itn_ref / cpu_s},
         note="working set (97 MB CSR + 7 x 8 MB vectors) fits the 126 MB L2: not an HBM roofline case")


def config3(ctx):
    import scipy.sparse as sp
    m = 215
    n = m ** 3
    note("config3: building the 7-pt operator and its transpose on device")
    A = DeviceCsr.convdiff3d(ctx, m, 0.5, build_transpose=True)
    ones = DeviceVector(ctx, n).fill(1.0)
    rhs = DeviceVector(ctx, n)
    A.spmv(ones, rhs)
    S = DeviceSolver(ctx, "bicgstab", A)
    ms, _ = time_device(ctx, S, lambda: S.setup_dev(rhs, abstol=0.0, reltol=0.0, matvec_max=10 ** 12), 100, 10)
    spmv_b = 12 * A.nnz + 4 * (n + 1) + 16 * n
    it_bytes = 2 * spmv_b + 120 * n
    note("config3: %.4f ms per Bi-CGSTAB iteration; timing A x / A^T x" % ms)
    # A^T x and A x stand-alone
    y = DeviceVector(ctx, n)
    res = {}
    for trans in (0, 1):
        for _ in range(3):
            A.spmv(ones, y, trans=bool(trans))
        ctx.sync()
        ctx.timer_start()
        for _ in range(20):
            A.spmv(ones, y, trans=bool(trans))
        t = ctx.timer_stop() / 20
        res["spmv_T" if trans else "spmv"] = {"ms": t, "GBs": spmv_b / t / 1e6, "frac_of_peak": spmv_b / t / 1e6 / peak()}
    note("config3: " + json.dumps(res))
    # to convergence, reltol 1e-8, zero guess, device-resident
    S.setup_dev(rhs, abstol=1e-8, reltol=1e-8, matvec_max=20000)      # bounded: never hours on a GPU box
    t0 = time.perf_counter()
    st = S.run(32)
    solve_s = time.perf_counter() - t0
    x = S.solution()
    note("config3: solve %.3f s, %d matvecs, resid %.3e, done=%d" % (solve_s, st.n_matvec, st.resid_norm, st.done))
    # CPU reference: bounded sample (6 iterations) on the same operator
    ip, ix, dv = kr.convdiff3d_csr(m)
    M = sp.csr_matrix((dv, ix, ip), shape=(n, n))
    rhs_h = M @ np.ones(n)
    lo, K = ref_classes()
    steps_cpu = 6
    t0 = time.perf_counter()
    if K:
        ref = K["bicgstab"](lo.LinearOperator(n, n, lambda v: M @ v), abstol=0.0, reltol=0.0)
        ref.solve(rhs_h, matvec_max=2 * steps_cpu)
        kind = "reference"
    else:
        kr.bicgstab_solve(lambda v: M @ v, rhs_h, abstol=0.0, reltol=0.0, matvec_max=2 * steps_cpu)
        kind = "port"
    cpu_s = time.perf_counter() - t0
    emit(config="configs[3]: Bi-CGSTAB fp64 on 7-pt convection-diffusion grid 215^3 (N=%d, nnz=%d)" % (n, A.nnz),
         metric="bicgstab_iters_per_s", value=1e3 / ms, unit="iters/s", ms_per_step=ms,
         achieved_GBs=it_bytes / ms / 1e6, frac_of_measured_peak=it_bytes / ms / 1e6 / peak(),
         algorithmic_bytes_per_step=it_bytes, spmv=res,
         solve_to_reltol_1e8={"seconds": solve_s, "nMatvec": int(st.n_matvec), "residNorm": st.resid_norm,
                        "err_inf": float(np.max(np.abs(x - 1.0)))},
         cpu_baseline={"kind": kind, "iters_per_s": steps_cpu / cpu_s, "sample": "%d iterations" % steps_cpu})


def main():
    ctx = Context(0)
    which = sys.argv[1:] or ["0", "2", "3"]
    for w in which:
        {"0": config0, "2": config2, "3": config3}[w](ctx)


if __name__ == "__main__":
    main()
