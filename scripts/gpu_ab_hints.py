#!/usr/bin/env python3
"""A/B: L2 eviction-priority hints on/off for the CG loop (config 2 and a smaller grid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pykrylov_b200.device import Context, DeviceCsr, DeviceSolver, DeviceVector
ctx = Context(0)
for g in (7071, 10000, 3162):
    n = g * g
    A = DeviceCsr.poisson2d(ctx, g)
    x = DeviceVector(ctx, n).fill(1.0)
    rhs = DeviceVector(ctx, n)
    A.spmv(x, rhs)
    bytes_it = 12 * A.nnz + 4 * (n + 1) + 16 * n + 72 * n
    for rep in range(2):
        for hints in (0, 1):
            ctx.set_option(1, hints)
            S = DeviceSolver(ctx, "cg", A)
            S.setup_dev(rhs, abstol=0.0, reltol=0.0, matvec_max=10 ** 9)
            S.iterate(10)
            ctx.sync()
            ctx.prof_enable(100)
            ctx.timer_start()
            S.iterate(100)
            ms = ctx.timer_stop() / 100
            k1 = ctx.prof_read()
            ctx.prof_enable(0)
            print("g=%5d hints=%d  CG iter %.4f ms  %7.1f GB/s  %8.1f it/s  K1 %.4f ms  resid %.17g"
                  % (g, hints, ms, bytes_it / ms / 1e6, 1e3 / ms, k1[1] / max(k1[0], 1), S.status().resid_norm), flush=True)
            del S
    del A, x, rhs
