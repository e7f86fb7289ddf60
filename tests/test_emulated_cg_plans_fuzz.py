"""Randomised interleavings of iterate / read / write calls on the three CG launch plans, on the
host emulation of the device logic (tests/emu).  The fused plans (KRY_OPT_CG_FUSE = 1, 2) defer
the p (and x) update of a trip into the next trip's SpMV launch and pay what they owe when a
vector is read or written from outside (cg_settle, the S_PSTATE / S_XPEND flags, the ping-pong
p buffers, the `fresh` launch variant).  Whatever the caller does in between -- reads in the
middle of a run, writes, convergence, a curvature exit, matvec_max -- the three plans must show
the same bits.  (Under emulation the inner products are summed in the same order for every
plan, so bit-identity is the right bar; on the GPU the same property is tested on fixed
sequences in tests/test_gpu_parity.py.)"""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle.csr_ref import CsrRef


def _matrix(rng, kind, n):
    if kind == "spd":
        B = sp.random(n, n, density=min(1.0, 6.0 / n), random_state=int(rng.integers(1 << 30)), format="csr")
        A = (B + B.T) * 0.5 + sp.identity(n) * (1.0 + abs(B).sum(axis=1).max())
    else:                                   # symmetric indefinite: the curvature test fires
        B = sp.random(n, n, density=min(1.0, 6.0 / n), random_state=int(rng.integers(1 << 30)), format="csr")
        A = (B + B.T) * 0.5 + sp.diags(rng.choice([-1.0, 1.0], size=n) * (1.0 + rng.random(n)))
    A = A.tocsr()
    A.sort_indices()
    return CsrRef.from_scipy(A)


def _run(dev, L, ctx, M, plan, script, rhs, guess, precon, params):
    ctx.set_option(L.KRY_OPT_CG_FUSE, plan)
    A = dev.DeviceCsr.from_arrays(ctx, M.shape, M.indptr, M.indices, M.data, symmetric=True)
    S = dev.DeviceSolver(ctx, "cg", A)
    d = np.abs(M.to_scipy().diagonal()) + 0.5
    S.set_precon_diag(d if precon else None, precon)
    S.setup(rhs, guess=guess, **params)
    out = []
    for op, arg in script:
        if op == "iterate":
            S.iterate(arg)
        elif op == "read":
            done = bool(S.status().done)
            out.append((arg, S.get_vector(arg) if arg != "solution" else S.solution(), done))
        elif op == "write":                  # transplant a vector mid-run (what the single-step tests do)
            S.set_vector(arg[0], arg[1])
        elif op == "status":
            st = S.status()
            out.append(("status", (st.done, st.definite, st.n_matvec, st.n_iter, st.hist_count, st.resid_norm,
                                   tuple(st.aux[:4]))))
    st = S.status()
    out.append(("final", (st.done, st.converged, st.definite, st.n_matvec, st.n_iter, st.resid_norm)))
    out.append(("hist", S.drain_history(st)))
    for name in ("solution", "r", "p"):
        out.append((name, S.get_vector(name) if name != "solution" else S.solution(), bool(st.done)))
    S._release()
    A._release()
    return out


def _same(a, b):
    if isinstance(a, np.ndarray):
        return np.array_equal(a, b, equal_nan=True)
    if isinstance(a, tuple):
        return len(a) == len(b) and all(_same(x, y) for x, y in zip(a, b))
    if isinstance(a, float) and a != a:
        return b != b
    return a == b


@pytest.mark.parametrize("seed", range(96))
def test_cg_plans_show_the_same_bits_under_any_interleaving(emu_ctx, seed):
    from pykrylov_b200 import _lib as L
    from pykrylov_b200 import device as dev
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(1, 90))
    kind = "spd" if seed % 4 else "indefinite"
    M = _matrix(rng, kind, n)
    rhs = M.matvec(rng.standard_normal(n)) if seed % 5 else np.zeros(n)
    guess = rng.standard_normal(n) if seed % 3 == 0 else None
    precon = int(rng.integers(0, 3))
    params = dict(abstol=0.0 if seed % 2 else 1e-8, reltol=0.0 if seed % 2 else 1e-6,
                  matvec_max=int(rng.integers(1, 4 * n + 4)), check_curvature=bool(seed % 7))
    script = []
    for _ in range(int(rng.integers(3, 14))):
        r = rng.random()
        if r < 0.45:
            script.append(("iterate", int(rng.integers(0, 9))))
        elif r < 0.8:
            script.append(("read", str(rng.choice(["solution", "x", "r", "p", "Ap"]))))
        elif r < 0.9:
            script.append(("status", None))
        else:
            script.append(("write", (str(rng.choice(["x", "r", "p"])), rng.standard_normal(n))))
    default = emu_ctx.get_option(L.KRY_OPT_CG_FUSE)
    try:
        runs = [_run(dev, L, emu_ctx, M, plan, script, rhs, guess, precon, params) for plan in (0, 1, 2)]
    finally:
        emu_ctx.set_option(L.KRY_OPT_CG_FUSE, default)
    for plan in (1, 2):
        assert len(runs[plan]) == len(runs[0])
        for k, (ra, rb) in enumerate(zip(runs[0], runs[plan])):
            # every observable -- p after `done` included: all plans perform the direction update of
            # the trip that met the stopping test, as cg.py:149-151 does
            assert ra[0] == rb[0] and _same(ra[1:], rb[1:]), (seed, plan, k, ra[0], script, params)
