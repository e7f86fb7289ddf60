#!/usr/bin/env python3
"""A few device-resident iterations of one solver at its BASELINE size, for ncu captures:
    python scripts/prof_solvers.py bicgstab|minres|tfqmr|cgs|cg [iterations]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pykrylov_b200 import _lib as L
from pykrylov_b200.device import Context, DeviceCsr, DeviceSolver, DeviceVector
from pykrylov_b200.gallery import kron_sym_jpwh

which = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
ctx = Context(0)
ctx.set_option(L.KRY_OPT_GRAPHS, 0)
if which == "minres":
    ctx.set_option(L.KRY_OPT_MINRES_PERSISTENT, 0)
    shape, ip, ix, dv = kron_sym_jpwh(os.path.join(ROOT, "tests", "golden", "jpwh_991.mtx"), 1009)
    A = DeviceCsr.from_arrays(ctx, shape, ip, ix, dv, symmetric=True)
    n = shape[0]
elif which == "bicgstab":
    A = DeviceCsr.convdiff3d(ctx, 215, 0.5)
    n = 215 ** 3
else:
    A = DeviceCsr.poisson2d(ctx, 3162)
    n = 3162 ** 2
ones = DeviceVector(ctx, n).fill(1.0)
rhs = DeviceVector(ctx, n)
A.spmv(ones, rhs)
S = DeviceSolver(ctx, which, A)
if which == "minres":
    S.setup_dev(rhs, abstol=0.0, reltol=0.0, matvec_max=10 ** 9, rtol=0.0, etol=0.0, window=5)
else:
    S.setup_dev(rhs, abstol=0.0, reltol=0.0, matvec_max=10 ** 12)
S.iterate(iters)
ctx.sync()
st = S.status()
print(which, "iterations", st.n_iter, "resid", st.resid_norm)
