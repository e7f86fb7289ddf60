#!/usr/bin/env python3
"""A short LSQR and SYMMLQ solve on the config-2 operator, launch by launch (no graph replay), for
`ncu --metrics gpu__time_duration.sum`: which launch of a trip costs what."""
import contextlib, io, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pykrylov_b200 import _lib as L
from pykrylov_b200.device import Context, DeviceCsr
from pykrylov_b200.linop import CsrLinearOperator
from pykrylov_b200.lls import LSQRFramework
from pykrylov_b200.symmlq import Symmlq

ctx = Context(0)
ctx.set_option(L.KRY_OPT_GRAPHS, 0)
g = int(sys.argv[1]) if len(sys.argv) > 1 else 3162
A = DeviceCsr.poisson2d(ctx, g)
op = CsrLinearOperator(A, symmetric=True)
rhs = np.ones(g * g)
ls = LSQRFramework(op, context=ctx)
ls.solve(rhs, itnlim=12, atol=0.0, btol=0.0, etol=0.0, conlim=1e300, show=False)
sy = Symmlq(op, context=ctx)
sy.solve(rhs, matvec_max=14, rtol=0.0)
print("lsqr itn", ls.itn, "symmlq matvecs", sy.nMatvec)
