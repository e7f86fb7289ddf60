import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import krylov_ref as kr
from oracle.csr_ref import load_mtx
from pykrylov_b200.device import Context, DeviceCsr, DeviceSolver
from test_gpu_parity import rel, srel, upload
ctx = Context(0)
M = load_mtx(os.path.join(ROOT, "tests/golden/jpwh_991.mtx")); n = 991
rng = np.random.default_rng(22)
rhs = M.matvec(rng.standard_normal(n)); guess = rng.standard_normal(n)
d = np.maximum(np.abs(M.to_scipy().diagonal()), 1.0)
print("diag range", np.abs(M.to_scipy().diagonal()).min(), np.abs(M.to_scipy().diagonal()).max())
A = upload(ctx, M)
st = kr.cgs_start(M, rhs, guess=guess.copy(), precon=lambda r: r / d, matvec_max=10**6)
S = DeviceSolver(ctx, "cgs", A); S.set_precon_diag(d, 2); S.setup(rhs, guess=guess, matvec_max=10**6)
print("thr", st.threshold, S.status().threshold)
for k in range(8):
    for v in ("x", "r", "r0", "u", "p"): S.set_vector(v, st[v])
    S.set_vector("y", st.p / d); S.set_scalar("rho", float(st.rho))
    kr.cgs_step(M, st); S.iterate(1); ds = S.status()
    cancel = abs(np.dot(st.r0, st.r)) / (np.linalg.norm(st.r0) * np.linalg.norm(st.r))
    print(k, "resid", st.residNorm, ds.resid_norm, "alpha", srel(S.get_scalar("alpha"), st.alpha),
          "beta", S.get_scalar("beta"), st.beta, "rho", S.get_scalar("rho"), st.rho, "cancel", cancel,
          "x", rel(S.get_vector("x"), st.x), "r", rel(S.get_vector("r"), st.r), "done", ds.done, "fin", st.finished)
