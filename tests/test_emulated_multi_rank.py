"""The row-sharded path (N > 1) on the host emulation of the device logic, any number of ranks.

Every rank is a thread of this process with its own emulated context.  csrc/comm.cu runs as it
is -- shard finalisation (need lists, boundary sets, [local | halo] remap), packed halo exchange,
the peer-memory inboxes -- with NCCL and CUDA IPC replaced by barriers between the threads
(tests/emu/emu_context.cpp); the solver loops are the library's own.  What the 2-GPU test of
tests/test_multi_gpu.py checks on hardware (and nothing can check at 4 or 8 ranks without the
box) is checked here for 2, 3, 4 and 8 ranks: bit-exact sharded SpMV, the oracle's CG
trajectory under every launch plan -- including the fused plans on shards with the updated
boundary entries travelling in the halo (KRY_OPT_CG_FUSE_SHARDS) -- with the in-kernel
all-reduce protocol and with the NCCL one, identical scalars on every rank, and consistent
mid-run reads."""
import ctypes as C
import threading

import numpy as np
import pytest

from oracle import krylov_ref as kr
from oracle.csr_ref import CsrRef


def run_ranks(world, worker, timeout=240):
    errors, out = [None] * world, [None] * world

    def body(rank):
        try:
            out[rank] = worker(rank)
        except BaseException as exc:                    # noqa: BLE001 -- reported by the main thread
            errors[rank] = exc

    threads = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout)
    assert not any(t.is_alive() for t in threads), "a rank is stuck in a collective"
    for r, e in enumerate(errors):
        if e is not None:
            raise AssertionError("rank %d: %r" % (r, e)) from e
    return out


@pytest.mark.parametrize("world,simt", [(2, 0), (3, 0), (4, 0), (8, 0), (2, 1), (3, 1)])
def test_sharded_cg_on_emulated_ranks(emu_ctx, world, simt):
    """simt = 1: the threads of a block are fibers and the genuine in-kernel all-reduce of
    common.cuh (stores into the peers' inboxes, sequence-number flags, spin, ordered sum) runs
    as written, each rank on its own host thread."""
    from pykrylov_b200 import _lib as L
    L.lib.kry_emu_set_fibers.restype = C.c_int
    L.lib.kry_emu_set_fibers.argtypes = [C.c_int]
    L.lib.kry_emu_set_fibers(simt)
    try:
        _sharded_cg(L, world)
    finally:
        L.lib.kry_emu_set_fibers(0)


def _sharded_cg(L, world):
    from pykrylov_b200 import device as dev
    from pykrylov_b200.comm import row_partition
    g = 20
    n = g * g
    ip, ix, dv = kr.poisson2d_csr(g)
    M = CsrRef((n, n), ip, ix, dv)
    x0 = np.random.default_rng(1).standard_normal(n)
    rhs = M.matvec(np.ones(n))
    r_ref = -rhs + M.matvec(x0)
    ref = kr.cg_solve(M, rhs)
    uid = C.create_string_buffer(L.KRY_COMM_ID_BYTES)
    L.call("kry_comm_unique_id", uid)
    ranges = row_partition(n, world)

    def worker(rank):
        lo, hi = ranges[rank]
        ctx = dev.Context(0)
        ctx.comm_init(world, rank, uid.raw)
        res = {}
        try:
            assert ctx.get_option(L.KRY_OPT_P2P) == 1               # the inboxes were mapped
            A = dev.DeviceCsr.poisson2d(ctx, g, lo, hi)
            A.shard_finalize(n, lo)
            S = dev.DeviceSolver(ctx, "cg", A)
            # sharded SpMV through the halo exchange is bit-exact per row
            S.setup(rhs[lo:hi], guess=x0[lo:hi], matvec_max=10 ** 6)
            assert np.array_equal(S.get_vector("r"), r_ref[lo:hi])
            assert abs(S.status().resid_norm0 - np.linalg.norm(r_ref)) <= 1e-12 * np.linalg.norm(r_ref)
            runs = {}
            for p2p in (1, 0):
                ctx.set_option(L.KRY_OPT_P2P, p2p)
                for fuse_shards, form in ((0, 0), (1, 1), (1, 2)):
                    ctx.set_option(L.KRY_OPT_CG_FUSE_SHARDS, fuse_shards)
                    ctx.set_option(L.KRY_OPT_CG_FUSE, form if fuse_shards else 2)
                    S.setup(rhs[lo:hi], matvec_max=2 * n)
                    st = S.run(5)
                    hist = S.drain_history(st)[:, 0]
                    assert st.n_matvec == ref.nMatvec, (p2p, fuse_shards, form, st.n_matvec, ref.nMatvec)
                    rh = np.array(ref.residHistory)
                    assert len(hist) == len(rh) and np.max(np.abs(hist - rh) / rh) <= 1e-9
                    xs = S.solution()
                    assert np.max(np.abs(xs - ref.x[lo:hi])) <= 1e-9
                    same = ctx.allgather_bytes(np.array([st.resid_norm]).tobytes())
                    assert len(set(same)) == 1                       # identical scalars on every rank
                    runs[(p2p, fuse_shards, form)] = (hist.copy(), xs.copy(), S.get_vector("r"), S.get_vector("p"))
            base = runs[(1, 0, 0)]
            for key, val in runs.items():                            # plans / reduce paths: same bits
                for a, b in zip(base, val):
                    assert np.array_equal(a, b), key
            # mid-run reads settle what the fused plan owes, then the run continues
            ctx.set_option(L.KRY_OPT_P2P, 1)
            ctx.set_option(L.KRY_OPT_CG_FUSE_SHARDS, 1)
            ctx.set_option(L.KRY_OPT_CG_FUSE, 2)
            S.setup(rhs[lo:hi], matvec_max=2 * n)
            S.iterate(5)
            x5, p5 = S.solution(), S.get_vector("p")
            S.iterate(4)
            ctx.set_option(L.KRY_OPT_CG_FUSE_SHARDS, 0)
            S0 = dev.DeviceSolver(ctx, "cg", A)
            S0.setup(rhs[lo:hi], matvec_max=2 * n)
            S0.iterate(5)
            assert np.array_equal(x5, S0.solution()) and np.array_equal(p5, S0.get_vector("p"))
            S0.iterate(4)
            assert np.array_equal(S.solution(), S0.solution()) and np.array_equal(S.get_vector("p"), S0.get_vector("p"))
            # the halo exchange rides in the SpMV launch (KRY_OPT_HALO_P2P): 2 launches per trip,
            # against pack kernel + ncclAllGather + 2 launches with the option off
            per_trip = {}
            for halo_p2p in (1, 0):
                ctx.set_option(L.KRY_OPT_CG_FUSE_SHARDS, 1)
                ctx.set_option(L.KRY_OPT_HALO_P2P, halo_p2p)
                S.setup(rhs[lo:hi], matvec_max=2 * n)
                S.iterate(2)
                l0 = ctx.launch_count()
                S.iterate(4)
                per_trip[halo_p2p] = (ctx.launch_count() - l0) / 4
            assert per_trip == {1: 2, 0: 3}, per_trip
            ctx.set_option(L.KRY_OPT_HALO_P2P, 1)
            res["max_send"] = A.shape[1] - A.shape[0]
            ctx.barrier()
        finally:
            ctx.close()
        return res

    out = run_ranks(world, worker)
    # the packed halo is one grid line per neighbour: interior ranks publish 2 g entries, so the
    # padded per-rank slot is g for two ranks and 2 g beyond, and the tail holds `world` slots
    assert all(o["max_send"] == world * g * (1 if world == 2 else 2) for o in out)


@pytest.mark.parametrize("method,world", [("bicgstab", 2), ("bicgstab", 5), ("cgs", 3), ("tfqmr", 4), ("minres", 3),
                                          ("minres", 8)])
def test_other_sharded_loops_follow_the_oracle(emu_ctx, method, world):
    """Bi-CGSTAB, CGS, TFQMR and MINRES on row shards (every SpMV input carries the halo tail,
    every fused inner product is all-reduced): the oracle's iteration count and history, the same
    solution, identical scalars on every rank -- with the in-kernel all-reduce protocol and with
    the NCCL one."""
    from pykrylov_b200 import _lib as L
    from pykrylov_b200 import device as dev
    from pykrylov_b200.comm import row_partition
    m = 7
    n = m ** 3
    if method == "minres":                       # symmetric: the 7-point Laplacian part only
        ip, ix, dv = kr.convdiff3d_csr(m, gamma=0.0)
    else:
        ip, ix, dv = kr.convdiff3d_csr(m)
    M = CsrRef((n, n), ip, ix, dv)
    rhs = M.matvec(np.linspace(1.0, 2.0, n))
    oracle = dict(bicgstab=kr.bicgstab_solve, cgs=kr.cgs_solve, tfqmr=kr.tfqmr_solve)
    if method == "minres":
        ref = kr.minres_solve(M, rhs)
    else:
        ref = oracle[method](M, rhs, reltol=1e-8, matvec_max=2 * n)
    uid = C.create_string_buffer(L.KRY_COMM_ID_BYTES)
    L.call("kry_comm_unique_id", uid)
    ranges = row_partition(n, world)

    def worker(rank):
        lo, hi = ranges[rank]
        ctx = dev.Context(0)
        ctx.comm_init(world, rank, uid.raw)
        try:
            A = dev.DeviceCsr.from_arrays(ctx, (hi - lo, n), ip[lo:hi + 1] - ip[lo], ix[ip[lo]:ip[hi]], dv[ip[lo]:ip[hi]],
                                          symmetric=(method == "minres"))
            A.shard_finalize(n, lo)
            S = dev.DeviceSolver(ctx, method, A)
            runs = []
            for p2p in (1, 0):
                ctx.set_option(L.KRY_OPT_P2P, p2p)
                if method == "minres":
                    S.setup(rhs[lo:hi], abstol=0.0, reltol=0.0, matvec_max=5 * n, rtol=1e-12, etol=1e-6, window=5)
                else:
                    S.setup(rhs[lo:hi], abstol=1e-8, reltol=1e-8, matvec_max=2 * n)
                st = S.run(6)
                hist = S.drain_history(st)[:, 0]
                if method == "minres":
                    assert (int(st.istop), int(st.n_iter)) == (ref.istop, ref.itn)
                else:
                    assert st.n_matvec == ref.nMatvec and bool(st.converged) == bool(ref.converged)
                rh = np.array(ref.residHistory, dtype=float)
                k = min(len(rh), 10)
                assert len(hist) == len(rh) and np.max(np.abs(hist[:k] - rh[:k]) / rh[:k]) <= 1e-9
                xs = S.solution()
                assert np.max(np.abs(xs - ref.x[lo:hi])) <= 1e-7 * np.max(np.abs(ref.x))
                same = ctx.allgather_bytes(np.array([st.resid_norm]).tobytes())
                assert len(set(same)) == 1
                runs.append((hist.copy(), xs.copy()))
            for a, b in zip(*runs):                      # the two all-reduce paths: same bits
                assert np.array_equal(a, b)
            ctx.barrier()
        finally:
            ctx.close()
        return True

    assert all(run_ranks(world, worker))


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_preconditioned_minres_follows_the_oracle(emu_ctx, world):
    """MINRES with a diagonal preconditioner, device-resident on row shards (the preconditioned
    y = M r2 is the SpMV input there and carries the halo tail)."""
    from pykrylov_b200 import _lib as L
    from pykrylov_b200 import device as dev
    from pykrylov_b200.comm import row_partition
    m = 6
    n = m ** 3
    ip, ix, dv = kr.convdiff3d_csr(m, gamma=0.0)
    M = CsrRef((n, n), ip, ix, dv)
    rhs = M.matvec(np.linspace(1.0, 2.0, n))
    d = 0.5 + np.random.default_rng(4).random(n)
    ref = kr.minres_solve(M, rhs, precon=lambda r: r / d)
    uid = C.create_string_buffer(L.KRY_COMM_ID_BYTES)
    L.call("kry_comm_unique_id", uid)
    ranges = row_partition(n, world)

    def worker(rank):
        lo, hi = ranges[rank]
        ctx = dev.Context(0)
        ctx.comm_init(world, rank, uid.raw)
        try:
            A = dev.DeviceCsr.from_arrays(ctx, (hi - lo, n), ip[lo:hi + 1] - ip[lo], ix[ip[lo]:ip[hi]], dv[ip[lo]:ip[hi]],
                                          symmetric=True)
            A.shard_finalize(n, lo)
            S = dev.DeviceSolver(ctx, "minres", A)
            S.set_precon_diag(d[lo:hi], 2)
            S.setup(rhs[lo:hi], abstol=0.0, reltol=0.0, matvec_max=5 * n, rtol=1e-12, etol=1e-6, window=5)
            st = S.run(5)
            hist = S.drain_history(st)[:, 0]
            assert (int(st.istop), int(st.n_iter)) == (ref.istop, ref.itn)
            rh = np.array(ref.residHistory, dtype=float)
            assert len(hist) == len(rh) and np.max(np.abs(hist[:10] - rh[:10]) / rh[:10]) <= 1e-9
            assert np.max(np.abs(S.solution() - ref.x[lo:hi])) <= 1e-7 * np.max(np.abs(ref.x))
            ctx.barrier()
        finally:
            ctx.close()
        return True

    assert all(run_ranks(world, worker))


@pytest.mark.parametrize("world", [2, 3])
def test_hardware_worker_sections_on_emulated_ranks(emu_ctx, world):
    """tests/multi_gpu_worker.py -- what `pytest -m gpu` runs on 2, 4 and 8 real GPUs -- section by
    section on emulated ranks: the hardware test's own logic is exercised in the CPU suite, and a
    rank that would hang in a collective on the box hangs (and is reported) here first."""
    from pykrylov_b200 import _lib as L
    from pykrylov_b200 import device as dev
    import multi_gpu_worker as W
    uid = C.create_string_buffer(L.KRY_COMM_ID_BYTES)
    L.call("kry_comm_unique_id", uid)

    def worker(rank):
        ctx = dev.Context(0)
        ctx.comm_init(world, rank, uid.raw)
        lines = []
        try:
            W.section_cg_stencil(ctx, rank, world, lines.append)
            W.section_cg_irregular(ctx, rank, world, lines.append)
            W.section_other_loops(ctx, rank, world, lines.append)
            W.section_public_api(ctx, rank, world, lines.append)
            ctx.barrier()
        finally:
            ctx.close()
        return lines

    out = run_ranks(world, worker, timeout=900)
    assert all(len(o) >= 8 for o in out)
