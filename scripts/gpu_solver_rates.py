#!/usr/bin/env python3
"""Device-resident iteration rate of every loop on the config-2 operator (5-pt Laplacian, 10^7 rows),
against the bytes its launch plan moves and the measured HBM copy peak."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pykrylov_b200.device import Context, DeviceCsr, DeviceSolver, DeviceVector

peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
ctx = Context(0)
g = 3162
n = g * g
A = DeviceCsr.poisson2d(ctx, g)
spmv = 12 * A.nnz + 4 * (n + 1) + 16 * n
ones = DeviceVector(ctx, n).fill(1.0)
rhs = DeviceVector(ctx, n)
A.spmv(ones, rhs)
# (SpMVs per iteration, vector bytes per row the plan moves, reference accounting per row)
plans = {"cg": (1, 56, 72), "bicgstab": (2, 104, 120), "cgs": (2, 112, None), "tfqmr": (2, 224, None), "minres": (1, 80, 96)}
out = {}
for name, (nsp, moved_row, ref_row) in plans.items():
    S = DeviceSolver(ctx, name, A)
    best = 1e9
    for _ in range(3):
        if name == "minres":
            S.setup_dev(rhs, abstol=0.0, reltol=0.0, matvec_max=10 ** 9, rtol=0.0, etol=0.0, window=5)
        else:
            S.setup_dev(rhs, abstol=0.0, reltol=0.0, matvec_max=10 ** 12)
        S.iterate(12)
        ctx.sync()
        ctx.timer_start()
        S.iterate(60)
        best = min(best, ctx.timer_stop() / 60)
        assert not S.status().done
    moved = nsp * spmv + moved_row * n
    out[name] = dict(ms_per_iteration=best, iters_per_s=1e3 / best, moved_bytes=moved, moved_GBs=moved / best / 1e6,
                     frac_of_measured_peak=moved / best / 1e6 / peak)
    if ref_row:
        out[name]["reference_bytes_GBs"] = (nsp * spmv + ref_row * n) / best / 1e6
    print("%-9s %.4f ms/iteration  %7.1f it/s  moved %.0f MB -> %.0f GB/s = %.3f of the measured copy peak"
          % (name, best, 1e3 / best, moved / 1e6, out[name]["moved_GBs"], out[name]["frac_of_measured_peak"]), flush=True)
    S._release()
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_solver_rates.json"), "w"), indent=1)
