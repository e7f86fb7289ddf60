#!/usr/bin/env python3
"""Tuning sweep (gpurun): fused SpMV+dot kernel variants and the full CG iteration on
BASELINE config 2.  Prints one line per variant; writes gpurun_out/tune.json."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pykrylov_b200 import _lib as L
from pykrylov_b200.device import Context, DeviceCsr, DeviceSolver, DeviceVector

ctx = Context(0)
g = int(sys.argv[1]) if len(sys.argv) > 1 else 3162
n = g * g
A = DeviceCsr.poisson2d(ctx, g)
spmv_bytes = 12 * A.nnz + 4 * (n + 1) + 16 * n
x = DeviceVector(ctx, n).fill(1.0)
y = DeviceVector(ctx, n)
rhs = DeviceVector(ctx, n)
A.spmv(x, rhs)
out = {}
variants = []
for kind, name in ((L.KRY_SPMV_ROW, "row"), (L.KRY_SPMV_ROWPF, "rowpf"), (L.KRY_SPMV_ROWPF2, "rowpf2"),
                   (L.KRY_SPMV_ROWB8, "rowb8"), (L.KRY_SPMV_ROWB4, "rowb4")):
    for bps in (8, 16, 32, 64):
        variants.append((kind, 0, bps * 32 if bps < 64 else 0, "%s/bps%d" % (name, bps)))
variants += [(L.KRY_SPMV_STREAM, 4096, 512, "stream/t4096/b512"), (L.KRY_SPMV_TMA, 2048, 512, "tma/t2048/b512")]
for kind, tile, thr, name in variants:
    A.set_kernel(kind, tile, thr)
    for _ in range(3):
        A.spmv_dot(x, y, [x], slot0=0)
    best = 1e9
    for _ in range(3):
        ctx.flush_l2()
        ctx.timer_start()
        for _ in range(20):
            A.spmv_dot(x, y, [x], slot0=0)
        best = min(best, ctx.timer_stop() / 20)
    S = DeviceSolver(ctx, "cg", A)
    S.setup_dev(rhs, abstol=0.0, reltol=0.0, matvec_max=10 ** 9)
    S.iterate(10)
    ctx.sync()
    ctx.timer_start()
    S.iterate(100)
    it_ms = ctx.timer_stop() / 100
    out[name] = dict(spmv_dot_ms=best, spmv_GBs=spmv_bytes / best / 1e6, cg_iter_ms=it_ms,
                     cg_GBs=(spmv_bytes + 72 * n) / it_ms / 1e6, check=float(ctx.scalars(0, 1)[0]))
    print("%-22s spmv+dot %.4f ms %7.1f GB/s | CG iter %.4f ms %7.1f GB/s %8.1f it/s" %
          (name, best, out[name]["spmv_GBs"], it_ms, out[name]["cg_GBs"], 1e3 / it_ms), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tune.json"), "w"), indent=1)
