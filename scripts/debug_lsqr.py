import os, sys
import numpy as np, scipy.sparse as sp
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from fake_bridge import FakeBridge
import pykrylov_b200._engine as eng
from pykrylov_b200.lls import LSQRFramework
from pykrylov_b200.linop import LinearOperator, linop_from_scipy
from pykrylov_b200.device import Context
ctx = Context(0)
R = sp.random(600, 200, density=0.03, random_state=7, format="csr"); R.sort_indices()
b = np.random.default_rng(7).standard_normal(600)
dev = LSQRFramework(linop_from_scipy(R, context=ctx), context=ctx); dev.solve(b, store_resids=True)
Real = eng.HostBridge
eng.HostBridge = FakeBridge
cpu = LSQRFramework(LinearOperator(200, 600, lambda v: R @ v, matvec_transp=lambda u: R.T @ u)); cpu.solve(b, store_resids=True)
eng.HostBridge = Real
for k in ("istop", "itn", "r1norm", "Anorm", "Acond", "Arnorm", "xnorm"):
    print(k, getattr(dev, k), getattr(cpu, k))
for i, (a, c, d, e) in enumerate(zip(dev.resids, cpu.resids, dev.normal_eqns_resids, cpu.normal_eqns_resids)):
    print(i, abs(a - c) / c, abs(d - e) / e)
    if i > 6: break
# direct check of the transposed SpMV on the rectangular operator
x = np.random.default_rng(1).standard_normal(600)
op = linop_from_scipy(R, context=ctx)
print("T exact:", np.array_equal(op.T * x, R.T @ x), "fwd exact:", np.array_equal(op * x[:200], R @ x[:200]))
