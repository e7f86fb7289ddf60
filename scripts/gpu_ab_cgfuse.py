#!/usr/bin/env python3
"""A/B/C (gpurun, one GPU): CG launch plans KRY_OPT_CG_FUSE = 0 / 1 / 2 on BASELINE config 2
(grid 3162^2) and on the 10^8-row grid, each with the L2-hint variants.  The residual norm
after the timed iterations must be the same bits for every plan.  Writes
gpurun_out/ab_cgfuse.json."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pykrylov_b200 import _lib as L                                            # noqa: E402
from pykrylov_b200.device import Context, DeviceCsr, DeviceSolver, DeviceVector   # noqa: E402

ctx = Context(0)
grids = [int(a) for a in sys.argv[1:]] or [3162, 10000]
K = 200
out = {}
for g in grids:
    n = g * g
    A = DeviceCsr.poisson2d(ctx, g)
    ones = DeviceVector(ctx, n).fill(1.0)
    rhs = DeviceVector(ctx, n)
    A.spmv(ones, rhs)
    spmv_b = 12 * A.nnz + 4 * (n + 1) + 16 * n
    k1_bytes = {0: spmv_b, 1: spmv_b + 16 * n, 2: spmv_b + 32 * n}
    resid = {}
    for form in (0, 1, 2):
        for hints in (1, 5, 0):
            ctx.set_option(L.KRY_OPT_CG_FUSE, form)
            ctx.set_option(L.KRY_OPT_L2_HINTS, hints)
            S = DeviceSolver(ctx, "cg", A)
            best = 1e9
            for rep in range(3):
                S.setup_dev(rhs, abstol=0.0, reltol=0.0, matvec_max=10 ** 9)
                S.iterate(20)
                ctx.sync()
                ctx.timer_start()
                S.iterate(K)
                best = min(best, ctx.timer_stop() / K)
            st = S.status()
            assert st.n_iter == 20 + K and not st.done
            resid.setdefault("r", st.resid_norm)
            same = (st.resid_norm == resid["r"])
            # per-launch time of the SpMV launch (graphs are off while profiling)
            S.setup_dev(rhs, abstol=0.0, reltol=0.0, matvec_max=10 ** 9)
            S.iterate(20)
            ctx.sync()
            ctx.prof_enable(100)
            S.iterate(100)
            k1 = ctx.prof_read()
            ctx.prof_enable(0)
            k1_ms = k1[1] / max(k1[0], 1)
            key = "g%d/form%d/hints%d" % (g, form, hints)
            out[key] = dict(ms_per_iter=best, iters_per_s=1e3 / best, k1_ms=k1_ms,
                            k1_GBs=k1_bytes[form] / k1_ms / 1e6, k1_bytes=k1_bytes[form],
                            step_GBs_ref_bytes=(spmv_b + 72 * n) / best / 1e6,
                            resid=st.resid_norm, same_bits_as_form0=bool(same))
            print("%-22s %.4f ms/iter %8.1f it/s | K1 %.4f ms %7.1f GB/s | resid %.17g %s"
                  % (key, best, 1e3 / best, k1_ms, out[key]["k1_GBs"], st.resid_norm, "same" if same else "DIFFERENT"),
                  flush=True)
            S._release()
    ctx.set_option(L.KRY_OPT_CG_FUSE, 0)
    ctx.set_option(L.KRY_OPT_L2_HINTS, 1)
    del A, ones, rhs
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "ab_cgfuse.json"), "w"), indent=1)
