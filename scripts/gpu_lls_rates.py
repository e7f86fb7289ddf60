#!/usr/bin/env python3
"""Per-trip time of the device-resident lls / SYMMLQ loops through the public classes, with the trip
replayed as a CUDA graph (default) and enqueued launch by launch (KRY_OPT_GRAPHS = 0), on a small
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
for g, lo, hi in ((256, 60, 360), (3162, 24, 72)):
    n = g * g
    A = DeviceCsr.poisson2d(ctx, g)
    op = CsrLinearOperator(A, symmetric=True)
    rhs = np.ones(n)
    for name, cls in (("lsqr", LSQRFramework), ("lsmr", LSMRFramework), ("craig", CRAIGFramework),
                      ("craigmr", CRAIGMRFramework), ("symmlq", Symmlq)):
        row = {}
        for graphs in (1, 0):
            ctx.set_option(L.KRY_OPT_GRAPHS, graphs)
            best = 1e9
            for _ in range(3):
                t = []
                for cap in (lo, hi):
                    s = cls(op, context=ctx)
                    ctx.sync()
                    t0 = time.perf_counter()
                    with contextlib.redirect_stdout(io.StringIO()):      # CRAIG-MR prints every trip, like the reference
                        if name == "symmlq":
                            s.solve(rhs, matvec_max=2 * cap + 2, rtol=0.0)
                            done = (s.nMatvec - 2) // 2
                        else:
                            s.solve(rhs, itnlim=cap, atol=0.0, btol=0.0, etol=0.0, conlim=1e300, show=False)
                            done = s.itn
                    ctx.sync()
                    t.append((time.perf_counter() - t0, done))
                best = min(best, (t[1][0] - t[0][0]) / max(1, t[1][1] - t[0][1]))
            row["graph" if graphs else "enqueued"] = best * 1e6
        ctx.set_option(L.KRY_OPT_GRAPHS, 1)
        out["%s/g%d" % (name, g)] = row
        print("%-8s n=%9d  %8.1f us/trip as a graph   %8.1f us/trip enqueued   (%.2fx)"
              % (name, n, row["graph"], row["enqueued"], row["enqueued"] / row["graph"]), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2r_lls_rates.json"), "w"), indent=1)
