"""Randomised interleavings of iterate / read / write calls on the two MINRES launch plans
(KRY_OPT_MINRES_FUSE), on the host emulation of the device logic.  The 2-launch plan defers the
w / x update of a trip into the next trip's second launch and pays what it owes when a vector is
read or written from outside (minres_settle, M_WPEND, the `fresh` launch variant); the two plans
must show the same bits whatever the caller does in between."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle.csr_ref import CsrRef


def _run(dev, L, ctx, M, plan, script, rhs, params):
    ctx.set_option(L.KRY_OPT_MINRES_FUSE, plan)
    A = dev.DeviceCsr.from_arrays(ctx, M.shape, M.indptr, M.indices, M.data, symmetric=True)
    S = dev.DeviceSolver(ctx, "minres", A)
    S.setup(rhs, **params)
    out = []
    for op, arg in script:
        if op == "iterate":
            S.iterate(arg)
        elif op == "read":
            out.append((arg, S.get_vector(arg) if arg != "solution" else S.solution()))
        elif op == "write":
            S.set_vector(arg[0], arg[1])
        elif op == "status":
            st = S.status()
            out.append(("status", (st.done, st.istop, st.n_matvec, st.n_iter, st.hist_count, st.resid_norm,
                                   tuple(st.aux[:13]))))
    st = S.status()
    out.append(("final", (st.done, st.converged, st.istop, st.n_iter, st.resid_norm, tuple(st.aux[:13]))))
    out.append(("hist", S.drain_history(st)))
    for name in ("solution", "w", "w2", "r1", "r2"):
        out.append((name, S.get_vector(name) if name != "solution" else S.solution()))
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


@pytest.mark.parametrize("seed", range(64))
def test_minres_plans_show_the_same_bits_under_any_interleaving(emu_ctx, seed):
    from pykrylov_b200 import _lib as L
    from pykrylov_b200 import device as dev
    rng = np.random.default_rng(5000 + seed)
    n = int(rng.integers(2, 80))
    B = sp.random(n, n, density=min(1.0, 6.0 / n), random_state=int(rng.integers(1 << 30)), format="csr")
    A = ((B + B.T) * 0.5 + sp.diags(rng.choice([-1.0, 1.0], size=n) * (0.5 + rng.random(n)))).tocsr()
    A.sort_indices()
    M = CsrRef.from_scipy(A)
    rhs = M.matvec(rng.standard_normal(n)) if seed % 6 else np.zeros(n)
    params = dict(matvec_max=int(rng.integers(0, 3 * n + 3)), shift=float(rng.choice([0.0, 0.3])),
                  rtol=0.0 if seed % 2 else 1e-12, etol=0.0 if seed % 3 else 1e-6, window=int(rng.integers(1, 8)))
    script = []
    for _ in range(int(rng.integers(3, 14))):
        r = rng.random()
        if r < 0.45:
            script.append(("iterate", int(rng.integers(0, 9))))
        elif r < 0.8:
            script.append(("read", str(rng.choice(["solution", "x", "w", "w2", "r1", "r2"]))))
        elif r < 0.9:
            script.append(("status", None))
        else:
            script.append(("write", (str(rng.choice(["x", "w", "w2", "r1", "r2"])), rng.standard_normal(n))))
    default = emu_ctx.get_option(L.KRY_OPT_MINRES_FUSE)
    try:
        runs = [_run(dev, L, emu_ctx, M, plan, script, rhs, params) for plan in (0, 1)]
    finally:
        emu_ctx.set_option(L.KRY_OPT_MINRES_FUSE, default)
    assert len(runs[0]) == len(runs[1])
    for k, (ra, rb) in enumerate(zip(runs[0], runs[1])):
        assert ra[0] == rb[0] and _same(ra[1:], rb[1:]), (seed, k, ra[0], script, params)
