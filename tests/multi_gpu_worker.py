"""One rank of the hardware test of the row-sharded path (tests/test_multi_gpu.py launches
WORLD_SIZE of these, one per GPU; also runnable under torchrun).  Everything is compared with
the 1-process oracle (oracle/krylov_ref.py, oracle/csr_ref.c) on the same inputs:

  1. sharded SpMV + all-reduced dot, bit-exact per row (5-point stencil and an irregular matrix)
  2. CG to convergence under every launch plan, with the in-kernel NVLink all-reduce and with
     ncclAllReduce: iteration count, history, solution, identical scalars on every rank, and the
     same bits (history, x, r, p) under every plan / reduction path; mid-run reads
  3. CG on an irregular (non-stencil) SPD shard
  4. Bi-CGSTAB, CGS, TFQMR on the 7-point convection-diffusion operator, MINRES (plain and with
     a diagonal preconditioner) on its symmetric part, both reduction paths
  5. the public API on shards: CG / Minres classes with default keywords (symmetry check and
     iteration caps must be global decisions)

Prints "rank R ok" at the end; any assertion kills the rank with a traceback.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import krylov_ref as kr                                  # noqa: E402
from oracle.csr_ref import CsrRef                                    # noqa: E402
from pykrylov_b200 import _lib as L                                  # noqa: E402
from pykrylov_b200.comm import init_from_env, row_partition         # noqa: E402
from pykrylov_b200.device import DeviceCsr, DeviceSolver            # noqa: E402


def shard(ctx, ip, ix, dv, n, lo, hi, symmetric=False):
    A = DeviceCsr.from_arrays(ctx, (hi - lo, n), ip[lo:hi + 1] - ip[lo], ix[ip[lo]:ip[hi]], dv[ip[lo]:ip[hi]],
                              symmetric=symmetric)
    A.shard_finalize(n, lo)
    return A


def same_on_all_ranks(ctx, value):
    return len(set(ctx.allgather_bytes(np.array([value], dtype=np.float64).tobytes()))) == 1


def section_cg_stencil(ctx, rank, world, log):
    g = 200
    n = g * g
    lo, hi = row_partition(n, world)[rank]
    A = DeviceCsr.poisson2d(ctx, g, lo, hi)
    A.shard_finalize(n, lo)
    ip, ix, dv = kr.poisson2d_csr(g)
    M = CsrRef((n, n), ip, ix, dv)
    x = np.random.default_rng(1).standard_normal(n)
    S = DeviceSolver(ctx, "cg", A)
    rhs = M.matvec(np.ones(n))
    # the setup kernel of a guess-started CG computes r = A x - b through the halo exchange
    S.setup(rhs[lo:hi], guess=x[lo:hi], matvec_max=10 ** 6)
    r_ref = -rhs + M.matvec(x)
    assert np.array_equal(S.get_vector("r"), r_ref[lo:hi])
    st0 = S.status()
    assert abs(st0.resid_norm0 - np.linalg.norm(r_ref)) <= 1e-12 * np.linalg.norm(r_ref)
    ref = kr.cg_solve(M, rhs)
    rh = np.array(ref.residHistory)
    runs = {}
    for p2p in (1, 0):
        ctx.set_option(L.KRY_OPT_P2P, p2p)
        for halo_p2p in ((1, 0) if p2p else (0,)):
            ctx.set_option(L.KRY_OPT_HALO_P2P, halo_p2p)
            for fuse_shards, form in ((0, 0), (1, 1), (1, 2)):
                ctx.set_option(L.KRY_OPT_CG_FUSE_SHARDS, fuse_shards)
                ctx.set_option(L.KRY_OPT_CG_FUSE, form if fuse_shards else 2)
                S.setup(rhs[lo:hi], matvec_max=2 * n)
                st = S.run(16)
                hist = S.drain_history(st)[:, 0]
                key = (p2p, halo_p2p, fuse_shards, form)
                assert st.n_matvec == ref.nMatvec, (key, st.n_matvec, ref.nMatvec)
                assert len(hist) == len(rh) and np.max(np.abs(hist - rh) / rh) <= 1e-9, key
                xs = S.solution()
                assert np.max(np.abs(xs - ref.x[lo:hi])) <= 1e-9, key
                assert same_on_all_ranks(ctx, st.resid_norm), key
                runs[key] = (hist.copy(), xs.copy(), S.get_vector("r"), S.get_vector("p"))
    # Launch plans and halo paths only move work between launches: same bits.  The two all-reduce
    # paths agree to the bit on 2 ranks only (a + b commutes); beyond that ncclAllReduce sums in its
    # own tree/ring order, the in-kernel all-reduce in rank order -- each is deterministic and the same
    # on every rank, so plans are compared bit-for-bit inside a reduction path and to rounding across.
    for key, val in runs.items():
        base = runs[(key[0], key[0], 0, 0)]
        for name, a, b in zip(("hist", "x", "r", "p"), base, val):
            assert np.array_equal(a, b), (key, name)
    for name, a, b in zip(("hist", "x"), runs[(1, 1, 0, 0)], runs[(0, 0, 0, 0)]):
        if world == 2:
            assert np.array_equal(a, b), name
        assert np.max(np.abs(a - b)) <= 1e-9 * np.max(np.abs(a)), name
    log("cg stencil: %d plan/path combinations bit-identical, nmv=%d" % (len(runs), ref.nMatvec))
    # mid-run reads settle what the fused plan still owes, then the run continues
    ctx.set_option(L.KRY_OPT_P2P, 1)
    ctx.set_option(L.KRY_OPT_HALO_P2P, 1)
    ctx.set_option(L.KRY_OPT_CG_FUSE_SHARDS, 1)
    ctx.set_option(L.KRY_OPT_CG_FUSE, 2)
    S.setup(rhs[lo:hi], matvec_max=2 * n)
    S.iterate(5)
    x5, p5 = S.solution(), S.get_vector("p")
    S.iterate(4)
    ctx.set_option(L.KRY_OPT_CG_FUSE_SHARDS, 0)
    S0 = DeviceSolver(ctx, "cg", A)
    S0.setup(rhs[lo:hi], matvec_max=2 * n)
    S0.iterate(5)
    assert np.array_equal(x5, S0.solution()) and np.array_equal(p5, S0.get_vector("p"))
    S0.iterate(4)
    assert np.array_equal(S.solution(), S0.solution()) and np.array_equal(S.get_vector("p"), S0.get_vector("p"))
    ctx.set_option(L.KRY_OPT_CG_FUSE_SHARDS, 1)
    log("cg stencil: mid-run reads consistent")
    return A, M, rhs, (lo, hi)


def section_cg_irregular(ctx, rank, world, log):
    import scipy.sparse as sp
    n = 4001
    R = sp.random(n, n, density=0.002, random_state=11, format="csr")
    R = (R + R.T).tocsr()
    R = (R + sp.diags(np.asarray(abs(R).sum(axis=1)).ravel() + 1.0)).tocsr()      # SPD, irregular pattern
    R.sort_indices()
    ip, ix, dv = R.indptr, R.indices, R.data
    M = CsrRef((n, n), ip, ix, dv)
    lo, hi = row_partition(n, world)[rank]
    A = shard(ctx, ip, ix, dv, n, lo, hi, symmetric=True)
    x = np.random.default_rng(3).standard_normal(n)
    S = DeviceSolver(ctx, "cg", A)
    rhs = M.matvec(np.linspace(-1.0, 1.0, n))
    S.setup(rhs[lo:hi], guess=x[lo:hi], matvec_max=10 ** 6)
    assert np.array_equal(S.get_vector("r"), (-rhs + M.matvec(x))[lo:hi])          # SpMV bit-exact per row
    ref = kr.cg_solve(M, rhs)
    rh = np.array(ref.residHistory)
    runs = []
    for halo_p2p, form in ((1, 2), (0, 2), (1, 0)):
        ctx.set_option(L.KRY_OPT_HALO_P2P, halo_p2p)
        ctx.set_option(L.KRY_OPT_CG_FUSE, form if form else 2)
        ctx.set_option(L.KRY_OPT_CG_FUSE_SHARDS, 1 if form else 0)
        S.setup(rhs[lo:hi], matvec_max=2 * n)
        st = S.run(8)
        hist = S.drain_history(st)[:, 0]
        assert st.n_matvec == ref.nMatvec and len(hist) == len(rh)
        assert np.max(np.abs(hist - rh) / rh) <= 1e-9
        xs = S.solution()
        assert np.max(np.abs(xs - ref.x[lo:hi])) <= 1e-9 * np.max(np.abs(ref.x))
        runs.append((hist.copy(), xs.copy()))
    for a, b in zip(runs[0], runs[1]):
        assert np.array_equal(a, b)
    for a, b in zip(runs[0], runs[2]):
        assert np.array_equal(a, b)
    ctx.set_option(L.KRY_OPT_HALO_P2P, 1)
    ctx.set_option(L.KRY_OPT_CG_FUSE, 2)
    ctx.set_option(L.KRY_OPT_CG_FUSE_SHARDS, 1)
    log("cg irregular: nmv=%d, max_send=%d" % (ref.nMatvec, (A.shape[1] - A.shape[0]) // world))


def section_other_loops(ctx, rank, world, log):
    m = 12
    n = m ** 3
    lo, hi = row_partition(n, world)[rank]
    oracle = dict(bicgstab=kr.bicgstab_solve, cgs=kr.cgs_solve, tfqmr=kr.tfqmr_solve)
    for method in ("bicgstab", "cgs", "tfqmr", "minres", "minres_precon"):
        sym = method.startswith("minres")
        ip, ix, dv = kr.convdiff3d_csr(m, gamma=0.0 if sym else 0.5)
        M = CsrRef((n, n), ip, ix, dv)
        rhs = M.matvec(np.linspace(1.0, 2.0, n))
        d = 0.5 + np.random.default_rng(4).random(n)
        if method == "minres":
            ref = kr.minres_solve(M, rhs)
        elif method == "minres_precon":
            ref = kr.minres_solve(M, rhs, precon=lambda r: r / d)
        else:
            ref = oracle[method](M, rhs, reltol=1e-8, matvec_max=2 * n)
        A = shard(ctx, ip, ix, dv, n, lo, hi, symmetric=sym)
        S = DeviceSolver(ctx, "minres" if sym else method, A)
        if method == "minres_precon":
            S.set_precon_diag(d[lo:hi], 2)
        runs = []
        for p2p, halo_p2p in ((1, 1), (1, 0), (0, 0)):
            ctx.set_option(L.KRY_OPT_P2P, p2p)
            ctx.set_option(L.KRY_OPT_HALO_P2P, halo_p2p)
            if sym:
                S.setup(rhs[lo:hi], abstol=0.0, reltol=0.0, matvec_max=5 * n, rtol=1e-12, etol=1e-6, window=5)
            else:
                S.setup(rhs[lo:hi], abstol=1e-8, reltol=1e-8, matvec_max=2 * n)
            st = S.run(6)
            hist = S.drain_history(st)[:, 0]
            if sym:
                assert (int(st.istop), int(st.n_iter)) == (ref.istop, ref.itn), (method, st.istop, st.n_iter)
            else:
                assert st.n_matvec == ref.nMatvec and bool(st.converged) == bool(ref.converged), method
            rh = np.array(ref.residHistory, dtype=float)
            k = min(len(rh), 10)
            assert len(hist) == len(rh) and np.max(np.abs(hist[:k] - rh[:k]) / rh[:k]) <= 1e-9, method
            xs = S.solution()
            assert np.max(np.abs(xs - ref.x[lo:hi])) <= 1e-7 * np.max(np.abs(ref.x)), method
            assert same_on_all_ranks(ctx, st.resid_norm), method
            runs.append((hist.copy(), xs.copy()))
        for a, b in zip(runs[0], runs[1]):               # halo paths: same bits
            assert np.array_equal(a, b), method
        for a, b in zip(runs[0], runs[2]):               # reduction paths: same bits on 2 ranks, rounding beyond
            if world == 2:
                assert np.array_equal(a, b), method
            assert np.max(np.abs(a - b)) <= 1e-7 * np.max(np.abs(a)), method
        ctx.set_option(L.KRY_OPT_P2P, 1)
        ctx.set_option(L.KRY_OPT_HALO_P2P, 1)
        log("%s: follows the oracle (%d history entries), paths bit-identical" % (method, len(ref.residHistory)))


def section_public_api(ctx, rank, world, log):
    import io
    from contextlib import redirect_stdout
    from pykrylov_b200.cg import CG
    from pykrylov_b200.linop import CsrLinearOperator
    from pykrylov_b200.minres import Minres
    m = 9
    n = m ** 3                      # 729 rows: the local slices differ by one row for world = 2, 4, 8
    lo, hi = row_partition(n, world)[rank]
    ip, ix, dv = kr.convdiff3d_csr(m, gamma=0.0)
    M = CsrRef((n, n), ip, ix, dv)
    rhs = M.matvec(np.linspace(1.0, 2.0, n))
    op = CsrLinearOperator(shard(ctx, ip, ix, dv, n, lo, hi, symmetric=True))
    # Minres with its default keywords: check=True runs the randomised symmetry test, whose inner
    # products are partial sums on a shard; itnlim defaults to 5 n (global n)
    ref = kr.minres_solve(M, rhs)
    mr = Minres(op)
    with redirect_stdout(io.StringIO()):
        mr.solve(rhs[lo:hi])
    assert (mr.istop, mr.itn) == (ref.istop, ref.itn), (mr.istop, mr.itn, ref.istop, ref.itn)
    assert np.max(np.abs(mr.x - ref.x[lo:hi])) <= 1e-7 * np.max(np.abs(ref.x))
    # CG capped by its default matvec_max on a tiny cap: every rank stops at the same trip
    cref = kr.cg_solve(M, rhs)
    cg = CG(op)
    cg.solve(rhs[lo:hi], check_symmetric=True)
    assert cg.nMatvec == cref.nMatvec and cg.converged == cref.converged
    assert np.max(np.abs(cg.bestSolution - cref.x[lo:hi])) <= 1e-9
    log("public API on shards: Minres(check=True) istop %d itn %d; CG nmv %d" % (mr.istop, mr.itn, cg.nMatvec))


def main():
    ctx, rank, world = init_from_env()

    def log(msg):
        if rank == 0:
            print("[world %d] %s" % (world, msg), flush=True)

    log("p2p all-reduce %d, p2p halo %d" % (ctx.get_option(L.KRY_OPT_P2P), ctx.get_option(L.KRY_OPT_HALO_P2P)))
    section_cg_stencil(ctx, rank, world, log)
    section_cg_irregular(ctx, rank, world, log)
    section_other_loops(ctx, rank, world, log)
    section_public_api(ctx, rank, world, log)
    ctx.barrier()
    print("rank %d ok" % rank, flush=True)


if __name__ == "__main__":
    main()
