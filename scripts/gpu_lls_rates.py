#!/usr/bin/env python3
"""Per-trip time of the device-resident lls / SYMMLQ loops through the public classes: the default (fused
launches, trip replayed as a CUDA graph), the three-launch form of the products (KRY_LLS_FUSE=0) and
the trip enqueued launch by launch (KRY_OPT_GRAPHS = 0), on a small
(launch-bound) and the config-2 (bandwidth-bound) 5-point Laplacian.  Per-trip time = difference of
two solves with different iteration caps (set-up and download cancel), best of 3."""
import contextlib, io, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pykrylov_b200 import _lib as L
from pykrylov_b200.device import Context, DeviceCsr, DeviceVector
from pykrylov_b200.linop import CsrLinearOperator
from pykrylov_b200.lls import LSQRFramework, LSMRFramework, CRAIGFramework, CRAIGMRFramework
from pykrylov_b200.symmlq import Symmlq

ctx = Context(0)
out = {}
for g, lo, hi in ((256, 100, 1100), (3162, 40, 340)):
    n = g * g
    A = DeviceCsr.poisson2d(ctx, g)
    op = CsrLinearOperator(A, symmetric=True)
    rhs = np.ones(n)
    for name, cls in (("lsqr", LSQRFramework), ("lsmr", LSMRFramework), ("craig", CRAIGFramework),
                      ("craigmr", CRAIGMRFramework), ("symmlq", Symmlq)):
        row = {}
        if name == "symmlq" and g == 256:
            lo, hi = 50, 250                                      # SYMMLQ converges on this operator after 269 trips
        for graphs, fuse in ((1, 1), (1, 2), (1, 3), (1, 0), (0, 1)):
            ctx.set_option(L.KRY_OPT_GRAPHS, graphs)
            os.environ["KRY_LLS_FUSE"] = str(fuse)
            best = {lo: 1e9, hi: 1e9}
            trips = {}
            for rep in range(5):                                   # rep 0 warms up (transpose, pools, first-use)
                for cap in (lo, hi):
                    s = cls(op, context=ctx)
                    ctx.sync()
                    t0 = time.perf_counter()
                    with contextlib.redirect_stdout(io.StringIO()):      # CRAIG-MR prints every trip, like the reference
                        if name == "symmlq":
                            s.solve(rhs, matvec_max=cap + 2, rtol=0.0)
                            trips[cap] = s.nMatvec - 2                    # one product per trip
                        else:
                            ret = s.solve(rhs, itnlim=cap, atol=0.0, btol=0.0, etol=0.0, conlim=1e300, show=False)
                            trips[cap] = ret[2] if name == "lsmr" else s.itn     # the reference's LSMR returns itn
                    ctx.sync()
                    if rep:
                        best[cap] = min(best[cap], time.perf_counter() - t0)
            assert trips[hi] - trips[lo] >= (hi - lo) // 2, (name, trips)
            row[("graph" if graphs else "enqueued") + {1: "", 2: "_step_in_spmv", 3: "_step_apart", 0: "_3launch"}[fuse]] = (best[hi] - best[lo]) / (trips[hi] - trips[lo]) * 1e6
            row["trips"] = [trips[lo], trips[hi]]
        ctx.set_option(L.KRY_OPT_GRAPHS, 1)
        out["%s/g%d" % (name, g)] = row
        os.environ.pop("KRY_LLS_FUSE", None)
        print("%-8s n=%9d us/trip: %7.1f default | %7.1f phase inside the SpMV launch | %7.1f phase as its own launch | "
              "%7.1f three-launch form | %7.1f default, enqueued (no graph)"
              % (name, n, row["graph"], row["graph_step_in_spmv"], row["graph_step_apart"], row["graph_3launch"],
                 row["enqueued"]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2r_lls_rates.json"), "w"), indent=1)
