#!/usr/bin/env python3
"""bench.py -- the headline benchmark of BASELINE.json on B200.

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU arm

A *step* is one pass of the hot path: one CG iteration on the 5-point Laplacian -- two
launches on one GPU ([x, p updates of the previous trip] + CSR SpMV + p.Ap | r update +
r.r + scalar recurrence + stopping test), three on row shards.

  N = 1 : BASELINE.json configs[1] -- CG fp64, 5-pt Poisson, grid 3162^2 (N = 9 998 244).
  N > 1 : BASELINE.json configs[4] -- the row-sharded 10^8-row operator (grid 10000^2),
          1-D row blocks, packed-halo ncclAllGather + ncclAllReduce of the scalars
          (strong scaling: the total problem is fixed for every N >= 2; rank 0 also
          times the same operator on one GPU so the speed-up is in the same line).

`value` is CG iterations/s with everything resident in HBM (CUDA events, max over
ranks); `e2e` is the same metric through the public pykrylov-style API
(CG(op).solve(rhs) with host buffers: H2D of rhs, per-check status reads, D2H of x
inside the timed region).  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

G_CONFIG2 = 3162          # grid of configs[1]
G_CONFIG5 = 10000         # grid of configs[4]
METRIC = "cg_iters_per_s"
UNIT = "iters/s"


def spmv_bytes(n, nnz):
    """Algorithmic bytes of one CSR SpMV."""
    return 12 * nnz + 4 * (n + 1) + 16 * n


def cg_iter_bytes(n, nnz):
    """Algorithmic bytes of one reference CG iteration: SpMV + 72 N."""
    return spmv_bytes(n, nnz) + 72 * n


# bytes the SpMV launch of each CG launch plan (KRY_OPT_CG_FUSE) must move beyond the plain
# SpMV: form 1 also reads r and writes the new p (16 N), form 2 also reads and writes x (32 N)
K1_EXTRA_BYTES_PER_ROW = {0: 0, 1: 16, 2: 32}
K1_NAMES = {0: "spmv_row_kernel<1,GatherPlain,CgEpiAp,CgFinAp> (fused CSR SpMV + p.Ap)",
            1: "spmv_row_kernel<1,CgGatherDir,CgEpiFused<1,0>,CgFinApFused> (p = beta p - r fused into CSR SpMV + p.Ap)",
            2: "spmv_row_kernel<1,CgGatherDir,CgEpiFused<1,1>,CgFinApFused> (x += alpha p ; p = beta p - r fused "
               "into CSR SpMV + p.Ap)"}
# what one whole iteration moves under each plan, per row, beyond the SpMV itself
ITER_VECTOR_BYTES_PER_ROW = {0: 72, 1: 64, 2: 56}


def workload_config(g, world):
    """The `config` object -- a description of the WORKLOAD only, identical for our arm and the
    reference arm (what each implementation does with it is reported under `implementation`)."""
    from pykrylov_b200.comm import row_partition
    n = g * g
    nnz = 5 * n - 4 * g
    lo, hi = row_partition(n, world)[0]
    n_loc = hi - lo
    # stored entries of rank 0's rows [0, n_loc): 5 per row minus the missing up / left / right neighbours
    nnz_loc = nnz if world == 1 else 5 * n_loc - min(g, n_loc) - (n_loc + g - 1) // g - n_loc // g
    name = ("BASELINE.json configs[1]: CG fp64, 5-pt Poisson Laplacian (gallery), grid %d^2" % g) if world == 1 else \
           ("BASELINE.json configs[4]: CG fp64, row-sharded 5-pt Laplacian grid %d^2, %d contiguous row blocks" % (g, world))
    return {"workload": name, "rows": n, "nnz": nnz, "rows_per_gpu": n_loc,
            "operator": "CSR int32 indices / fp64 values, sorted columns", "rhs": "A*ones",
            "stopping": "abstol=reltol=0 so exactly K iterations run",
            "l2": "inputs larger than L2: %.2f GB touched per iteration per GPU vs 126 MB L2"
                  % (cg_iter_bytes(n_loc, nnz_loc) / 1e9),
            "algorithmic_bytes_per_step_per_gpu": cg_iter_bytes(n_loc, nnz_loc)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


# ------------------------------------------------------------------ clocks
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "25"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def wait_first(self, timeout_s):
        """Block until the first sample has arrived (the polling loop is then in steady state)."""
        t0 = time.time()
        while (self.proc is not None and self.proc.poll() is None and not self.lines
               and time.time() - t0 < timeout_s):
            time.sleep(0.01)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------- our arm
def build_problem(ctx, g, rank, world):
    """Local rows of the g x g 5-point Laplacian, generated in HBM; rhs = A * ones."""
    from pykrylov_b200.comm import row_partition
    from pykrylov_b200.device import DeviceCsr, DeviceVector
    n = g * g
    lo, hi = row_partition(n, world)[rank]
    A = DeviceCsr.poisson2d(ctx, g, lo, hi)
    if world > 1:
        A.shard_finalize(n, lo)
    ones = A.input_vector()
    ones.fill(1.0)
    rhs = DeviceVector(ctx, hi - lo)
    A.spmv(ones, rhs)
    return A, rhs, n, lo, hi


def time_device_resident(ctx, A, rhs, steps, warmup, profile=True):
    """K iterations, inputs resident, CUDA events on the launching stream."""
    from pykrylov_b200.device import DeviceSolver
    S = DeviceSolver(ctx, "cg", A)
    if not profile:
        # One-off initialisation, outside every timed region: a throwaway solve long enough that the
        # loop captures and instantiates the CUDA graph it replays (one GPU; the graph is kept across
        # kry_solver_setup).  Without it the first cudaGraphInstantiate of the process lands inside the
        # timed K steps whenever W < 18 (measured on a fresh box at --steps 20 --warmup 5: +72 ms).
        # Sharded runs replay no graph; the same throwaway solve touches every code path, peer mapping
        # and NCCL channel once before the W warm-up steps.
        S.setup_dev(rhs, abstol=0.0, reltol=0.0, matvec_max=10 ** 12)
        S.iterate(36)
        ctx.sync()
    S.setup_dev(rhs, abstol=0.0, reltol=0.0, matvec_max=10 ** 12)
    S.iterate(warmup)
    ctx.sync()
    if profile:
        ctx.prof_enable(steps)
    if ctx.nranks > 1:
        ctx.barrier()
    ctx.sync()
    l0 = ctx.launch_count()
    t_host = time.perf_counter()
    ctx.timer_start()
    S.iterate(steps)
    t_enqueue = time.perf_counter() - t_host
    ms = ctx.timer_stop()                     # synchronises
    if ctx.rank == 0:
        print("[bench] %d iterations%s: enqueued in %.2f ms of host time, %.3f ms on the device"
              % (steps, " (per-launch events)" if profile else "", 1e3 * t_enqueue, ms), file=sys.stderr)
    launches = ctx.launch_count() - l0
    if ctx.nranks > 1:
        ctx.barrier()
        ms = float(ctx.allreduce([ms], op="max")[0])
    prof = ctx.prof_read() if profile else (0, 0.0)
    if profile:
        ctx.prof_enable(0)
    st = S.status()
    assert st.n_iter == warmup + steps and not st.done, "timed region did not run exactly K iterations"
    return ms, launches, prof, S, st


def time_e2e(ctx, op, rhs_host, steps, warmup_calls=1):
    """Public API, host buffers: CG(op).solve(rhs) -- H2D, K iterations with a status
    read per check interval, D2H of the solution, all inside the timed region."""
    from pykrylov_b200.cg import CG
    interval = 25
    for _ in range(warmup_calls):
        CG(op, abstol=0.0, reltol=0.0, check_interval=interval).solve(rhs_host, matvec_max=min(steps, 10))
    ctx.sync()
    if ctx.nranks > 1:
        ctx.barrier()
    t0 = time.perf_counter()
    cg = CG(op, abstol=0.0, reltol=0.0, check_interval=interval)
    cg.solve(rhs_host, matvec_max=steps)
    ctx.sync()
    dt = time.perf_counter() - t0
    if ctx.nranks > 1:
        dt = float(ctx.allreduce([dt], op="max")[0])
    assert cg.nMatvec == steps
    n_loc = rhs_host.shape[0]
    checks = (steps + interval - 1) // interval + 1
    d2h = 8 * n_loc + checks * 232 + (steps + 1) * 16       # x + status blocks + history
    return dt, 8 * n_loc / steps, d2h / steps, float(cg.residNorm)


def cpu_baseline(g, budget_s=20.0):
    """The reference's CPU path on the host cores, bounded sample of the same workload."""
    n = g * g
    t0 = time.perf_counter()
    res = run_reference_cg(g, max_steps=10 ** 9, warmup=1, budget_s=budget_s)
    res["wall_s_including_setup"] = time.perf_counter() - t0
    res["rows"] = n
    return res


def host_laplacian(g, block_rows=2000000):
    """scipy CSR (int32 indices) of the g x g 5-point Laplacian, assembled block-wise so that the
    10^8-row operator of configs[4] (6.4 GB of CSR) is never held twice."""
    import scipy.sparse as sp
    from oracle import krylov_ref as kr
    n = g * g
    nnz = 5 * n - 4 * g
    indptr = np.empty(n + 1, dtype=np.int32)
    indices = np.empty(nnz, dtype=np.int32)
    data = np.empty(nnz, dtype=np.float64)
    indptr[0] = 0
    at = 0
    for lo in range(0, n, block_rows):
        hi = min(n, lo + block_rows)
        ip, ix, dv = kr.poisson2d_csr(g, lo, hi)
        indptr[lo + 1:hi + 1] = ip[1:] + at
        indices[at:at + len(ix)] = ix
        data[at:at + len(dv)] = dv
        at += len(dv)
    assert at == nnz
    return sp.csr_matrix((data, indices, indptr), shape=(n, n))


def run_reference_cg(g, max_steps, warmup, budget_s):
    """CG of the reference (oracle/_ref if present, else the oracle port) on the
    g x g Laplacian with the scipy-CSR stand-in operator; returns iterations/s."""
    from oracle import krylov_ref as kr
    n = g * g
    M = host_laplacian(g)
    rhs = M @ np.ones(n)
    kind = "port"
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    CG = LinearOperator = None
    if os.path.isdir(os.path.join(ref_dir, "refpykrylov")):
        try:
            sys.path.insert(0, ref_dir)
            from refpykrylov.cg import CG
            from refpykrylov.linop import LinearOperator
            kind = "reference"
        except Exception:
            kind = "port"
    # one calibration iteration to size the sample
    t0 = time.perf_counter()
    y = M @ rhs
    float(np.dot(rhs, y))
    t_cal = max(time.perf_counter() - t0, 1e-3) * 2.4       # SpMV is ~42% of an iteration
    del y
    steps = int(max(3, min(max_steps, budget_s / t_cal)))
    warmup = int(max(1, min(warmup, 0.25 * budget_s / t_cal)))
    if kind == "reference":
        op = LinearOperator(n, n, lambda v: M @ v, symmetric=True)
        CG(op, abstol=0.0, reltol=0.0).solve(rhs, matvec_max=warmup)
        t0 = time.perf_counter()
        cg = CG(op, abstol=0.0, reltol=0.0)
        cg.solve(rhs, matvec_max=steps)
        dt = time.perf_counter() - t0
        done = cg.nMatvec
    else:
        kr.cg_solve(lambda v: M @ v, rhs, abstol=0.0, reltol=0.0, matvec_max=warmup)
        t0 = time.perf_counter()
        st = kr.cg_solve(lambda v: M @ v, rhs, abstol=0.0, reltol=0.0, matvec_max=steps)
        dt = time.perf_counter() - t0
        done = st.nMatvec
    try:
        from threadpoolctl import threadpool_info
        blas_threads = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        blas_threads = os.cpu_count() or 1
    return {"value": done / dt, "unit": UNIT, "cores": int(blas_threads), "kind": kind, "steps": int(done),
            "warmup": warmup, "ms_per_step": 1e3 * dt / done,
            "sample": "%d CG iterations (after %d warm-up) of pykrylov CG.solve (%s) on the full %dx%d 5-pt "
                      "Laplacian (%d rows), scipy-CSR operator; every host thread the path can use: SpMV and "
                      "AXPYs are single-threaded NumPy/SciPy, np.dot runs on %d OpenBLAS threads (`cores`); "
                      "host has %d logical cores" % (done, warmup, kind, g, g, n, blas_threads, os.cpu_count() or 0)}


def _time_iterations(ctx, S, setup, steps, warmup, reps=3):
    """Best of `reps`: `steps` device-resident iterations after `warmup`, CUDA events."""
    best = float("inf")
    for _ in range(reps):
        setup()
        S.iterate(warmup)
        ctx.sync()
        ctx.timer_start()
        S.iterate(steps)
        best = min(best, ctx.timer_stop() / steps)
        assert not S.status().done, "the timed iterations were not all live"
    return best


def other_config_minres(ctx, peak, peak_src):
    """BASELINE.json configs[2]: MINRES on kron(I_1009, sym(jpwh_991)), N = 999 919."""
    from pykrylov_b200.device import DeviceCsr, DeviceSolver
    from pykrylov_b200.gallery import kron_sym_jpwh
    shape, ip, ix, dv = kron_sym_jpwh(os.path.join(ROOT, "tests", "golden", "jpwh_991.mtx"), 1009)
    n, nnz = shape[0], len(dv)
    A = DeviceCsr.from_arrays(ctx, shape, ip, ix, dv, symmetric=True)
    ones = A.input_vector()
    ones.fill(1.0)
    from pykrylov_b200.device import DeviceVector
    rhs = DeviceVector(ctx, n)
    A.spmv(ones, rhs)
    S = DeviceSolver(ctx, "minres", A)
    # (t1 <= 1 fires after ~90 trips on this system: time inside the live range)
    ms = _time_iterations(ctx, S, lambda: S.setup_dev(rhs, abstol=0.0, reltol=0.0, matvec_max=10 ** 9, rtol=0.0,
                                                      etol=0.0, window=5), steps=48, warmup=16)
    it_bytes = spmv_bytes(n, nnz) + 96 * n
    S.setup_dev(rhs, abstol=0.0, reltol=0.0, matvec_max=5 * n, rtol=1e-12, etol=1e-6, window=5)
    st = S.run(16)
    out = {"workload": "BASELINE.json configs[2]: MINRES fp64 on kron(I_1009, sym(jpwh_991))", "rows": n, "nnz": nnz,
           "metric": "minres_iters_per_s", "value": 1e3 / ms, "unit": UNIT, "ms_per_step": ms,
           "to_convergence": {"istop": int(st.istop), "itn": int(st.n_iter), "rnorm": st.resid_norm},
           "roofline": {"bound": "hbm", "kernel": "whole MINRES iteration (SpMV + 96 N bytes, SURVEY.md 8d)",
                        "achieved": it_bytes / ms / 1e6, "peak": peak, "unit": "GB/s",
                        "frac": it_bytes / ms / 1e6 / peak, "traffic": None, "peak_source": peak_src,
                        "algorithmic_bytes_per_step": it_bytes,
                        "note": "working set 153 MB ~ L2: two dependent launches of ~25 us, latency-bound"}}
    S._release()
    A._release()
    return out


def other_config_bicgstab(ctx, peak, peak_src):
    """BASELINE.json configs[3]: Bi-CGSTAB on the 7-pt convection-diffusion operator, grid 215^3."""
    from pykrylov_b200.device import DeviceCsr, DeviceSolver, DeviceVector
    m = 215
    n = m ** 3
    A = DeviceCsr.convdiff3d(ctx, m, 0.5, build_transpose=True)
    ones = DeviceVector(ctx, n).fill(1.0)
    rhs = DeviceVector(ctx, n)
    A.spmv(ones, rhs)
    S = DeviceSolver(ctx, "bicgstab", A)
    ms = _time_iterations(ctx, S, lambda: S.setup_dev(rhs, abstol=0.0, reltol=0.0, matvec_max=10 ** 12),
                          steps=60, warmup=6, reps=2)
    sp_b = spmv_bytes(n, A.nnz)
    it_bytes = 2 * sp_b + 120 * n
    y = DeviceVector(ctx, n)
    spmv = {}
    for trans, key in ((False, "A_x"), (True, "AT_x")):
        for _ in range(3):
            A.spmv(ones, y, trans=trans)
        ctx.sync()
        ctx.timer_start()
        for _ in range(20):
            A.spmv(ones, y, trans=trans)
        t = ctx.timer_stop() / 20
        spmv[key] = {"ms": t, "GBs": sp_b / t / 1e6, "frac": sp_b / t / 1e6 / peak}
    out = {"workload": "BASELINE.json configs[3]: Bi-CGSTAB fp64, 7-pt convection-diffusion grid 215^3", "rows": n,
           "nnz": int(A.nnz), "metric": "bicgstab_iters_per_s", "value": 1e3 / ms, "unit": UNIT, "ms_per_step": ms,
           "spmv": spmv,
           "roofline": {"bound": "hbm", "kernel": "whole Bi-CGSTAB iteration (2 SpMV + 120 N bytes, SURVEY.md 8d; the "
                                                  "launch plan moves 2 SpMV + 104 N)",
                        "achieved": it_bytes / ms / 1e6, "peak": peak, "unit": "GB/s",
                        "frac": it_bytes / ms / 1e6 / peak, "traffic": None, "peak_source": peak_src,
                        "algorithmic_bytes_per_step": it_bytes,
                        "moved_bytes_per_step": 2 * sp_b + 104 * n,
                        "frac_of_moved_bytes": (2 * sp_b + 104 * n) / ms / 1e6 / peak}}
    S._release()
    A._release()
    return out


def load_traffic():
    """ncu-measured DRAM bytes per launch of the dominant kernel (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "spmv_dot_traffic.json")) as fh:
            return json.load(fh)
    except Exception:
        return None


class stdout_to_stderr(object):
    """Route the process's fd 1 to fd 2 for a while: NCCL prints its version banner on stdout
    when the communicator is created (NCCL_DEBUG=VERSION on some boxes), and stdout must carry
    exactly one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def main_ours(args):
    from pykrylov_b200.comm import init_from_env
    from pykrylov_b200.linop import CsrLinearOperator
    with stdout_to_stderr():
        ctx, rank, world = init_from_env()
    if world != args.gpus and rank == 0:
        print("warning: --gpus %d but WORLD_SIZE=%d" % (args.gpus, world), file=sys.stderr)
    g = args.grid or (G_CONFIG2 if world == 1 else G_CONFIG5)
    if world > 1 and args.nccl_allreduce:
        ctx.set_option(3, 0)
    if world > 1 and args.halo_p2p is not None:
        ctx.set_option(9, args.halo_p2p)
    fused_allreduce = bool(world > 1 and ctx.get_option(3))
    fused_halo = bool(fused_allreduce and ctx.get_option(9))
    sampler = ClockSampler(ctx.device)
    if rank == 0:
        sampler.start()          # early: nvidia-smi's start-up (NVML init) must not overlap a timed region
    A, rhs, n, lo, hi = build_problem(ctx, g, rank, world)
    nnz_total = 5 * n - 4 * g
    if rank == 0:
        sampler.wait_first(10.0)       # NVML start-up enumerates every GPU of the box: seconds on 8 GPUs
    if args.cg_fuse is not None:
        ctx.set_option(4, args.cg_fuse)
    if args.cg_fuse_shards is not None:
        ctx.set_option(5, args.cg_fuse_shards)
    form = ctx.get_option(4) if (world == 1 or ctx.get_option(5)) else 0    # launch plan in effect
    # timed region 1 -> `value`: K iterations exactly as a solve enqueues them (CUDA-graph replays
    # on one GPU), nothing else on the stream
    ms, launches, _, S, st = time_device_resident(ctx, A, rhs, args.steps, args.warmup, profile=False)
    S._release()
    # timed region 2 -> `roofline`: the same K iterations with a CUDA-event pair around every SpMV
    # launch (the events serialise the launches and switch graph replay off, so this region is a
    # few % slower than region 1; its own ms/step is reported next to the kernel time)
    ms_prof, _, prof, S2, _ = time_device_resident(ctx, A, rhs, args.steps, args.warmup, profile=True)
    S2._release()
    value = args.steps / (ms / 1e3)
    if os.environ.get("KRY_HALO_TRACE") and world > 1:
        t = ctx.halo_trace()
        print("[halo trace rank %d] launches %d: publish-entry %.1f us | waiting warps/launch %.0f: spin %.2f us, "
              "fence %.2f us, wait starts %.1f us after entry, latest end %.1f us after entry"
              % (rank, t[1], t[0] / max(t[1], 1) / 1e3, t[4] / max(t[1], 1), t[2] / max(t[4], 1) / 1e3,
                 t[3] / max(t[4], 1) / 1e3, t[6] / max(t[4], 1) / 1e3, t[5] / 1e3), file=sys.stderr)

    # roofline of the dominant kernel (fused SpMV+dot), from per-launch CUDA events
    peak, peak_src = measured_peak()
    n_loc, nnz_loc = hi - lo, A.nnz
    k1_ms = prof[1] / max(prof[0], 1)
    k1_bytes = spmv_bytes(n_loc, nnz_loc) + K1_EXTRA_BYTES_PER_ROW[form] * n_loc
    moved_per_step = spmv_bytes(n_loc, nnz_loc) + ITER_VECTOR_BYTES_PER_ROW[form] * n_loc
    achieved = k1_bytes / (k1_ms * 1e-3) / 1e9 if prof[0] else None
    traffic = load_traffic()
    traffic_ok = traffic is not None and world == 1 and g == G_CONFIG2 and traffic.get("cg_fuse", 0) == form
    roofline = {"bound": "hbm", "kernel": K1_NAMES[form], "cg_launch_plan": form,
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None,
                "traffic": traffic.get("dram_bytes_per_launch") if traffic_ok else None,
                "algorithmic_bytes_per_launch": k1_bytes,
                "avg_launch_ms": k1_ms, "launches_timed": prof[0], "peak_source": peak_src,
                # whole iteration: bytes this launch plan really has to move / time ...
                "step_achieved_GBs": moved_per_step * args.steps / (ms * 1e-3) / 1e9,
                # ... and the reference iteration's bytes (SpMV + 72 N) / time: exceeds what the
                # fused plans move, so it is an *effective* rate and may pass the HBM peak
                "step_effective_GBs_reference_bytes": cg_iter_bytes(n_loc, nnz_loc) * args.steps / (ms * 1e-3) / 1e9,
                "instrumented_region_ms_per_step": ms_prof / args.steps,
                "kernel_share_of_step": (prof[1] / ms_prof) if prof[0] else None}

    # end to end through the public API with host buffers
    rhs_host = ctx.pinned_array(n_loc)
    rhs.download(rhs_host)
    op = CsrLinearOperator(A)
    e2e_dt, h2d, d2h, _ = time_e2e(ctx, op, rhs_host, args.steps)
    clocks = sampler.stop() if rank == 0 else None        # sampled across all three timed regions
    e2e = {"value": args.steps / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "api": "pykrylov_b200.cg.CG(op, abstol=0, reltol=0).solve(rhs_host_pinned, matvec_max=K)",
           "wall_ms": 1e3 * e2e_dt}

    if world > 1:
        # the same launch on every rank (with the in-kernel all-reduce they wait for each other inside it;
        # with --nccl-allreduce these are the ranks' own kernel times: the skew the all-reduce absorbs)
        per_rank = [float(np.frombuffer(b, dtype=np.float64)[0])
                    for b in ctx.allgather_bytes(np.array([k1_ms], dtype=np.float64).tobytes())]
        roofline["avg_launch_ms_per_rank"] = per_rank

    extra = {}
    if world > 1 and not args.no_single:
        # same 10^8-row operator on ONE GPU (rank 0), for the speed-up claim
        ctx.barrier()
        if rank == 0:
            from pykrylov_b200.device import Context
            solo = Context(ctx.device)
            A1, rhs1, _, _, _ = build_problem(solo, g, 0, 1)
            # the same K iterations after the same W: its residual norm shows the 1 -> N trajectory agreement
            ms1, _, _, _, st1 = time_device_resident(solo, A1, rhs1, args.steps, args.warmup, profile=False)
            one = args.steps / (ms1 / 1e3)
            extra["one_gpu_same_workload"] = {"value": one, "unit": UNIT, "ms_per_step": ms1 / args.steps,
                                              "resid_norm_after_timed_region": st1.resid_norm}
            extra["speedup_vs_one_gpu"] = value / one
            extra["resid_rel_diff_vs_one_gpu"] = abs(st.resid_norm - st1.resid_norm) / st1.resid_norm
            solo.close()
        ctx.barrier()

    if world == 1 and g == G_CONFIG2 and not args.no_config5:
        # the strong-scaling base: the 10^8-row operator of configs[4] on this one GPU, so that the
        # N = 2/4/8 lines can be compared with a number from the same series of runs
        try:
            A5, rhs5, n5, _, _ = build_problem(ctx, G_CONFIG5, 0, 1)
            k5 = max(10, args.steps // 4)
            ms5, _, _, S5, _ = time_device_resident(ctx, A5, rhs5, k5, 3, profile=False)
            extra["config5_one_gpu"] = {"value": k5 / (ms5 / 1e3), "unit": UNIT, "steps": k5, "rows": n5,
                                        "workload": "BASELINE.json configs[4] operator (grid %d^2) on one GPU"
                                                    % G_CONFIG5}
            S5._release()
            A5._release()
            del rhs5
        except Exception as exc:                      # never lose the headline line over the side leg
            extra["config5_one_gpu"] = {"error": "%s: %s" % (type(exc).__name__, exc)}

    if world == 1 and g == G_CONFIG2 and not args.no_other_configs:
        # BASELINE.json configs[2] and configs[3] on this GPU, each with its own roofline object
        # (whole-iteration bytes of SURVEY.md 8d over the device-timed iteration), so that these
        # numbers are driver-visible too.  Never lose the headline line over a side leg.
        for name, leg in (("config3_minres", other_config_minres), ("config4_bicgstab", other_config_bicgstab)):
            try:
                extra[name] = leg(ctx, peak, peak_src)
            except Exception as exc:
                extra[name] = {"error": "%s: %s" % (type(exc).__name__, exc)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(g)

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic",
               "config": workload_config(g, world),
               "implementation": ("operator generated in HBM (kry_csr_create_poisson2d); 2 launches per iteration; "
                                  + ("halo: %s; scalars: %s" % (
                                      "boundary entries stored into the peers' halo tails over NVLink peer memory "
                                      "from inside the SpMV launch" if fused_halo else "pack kernel + one ncclAllGather",
                                      "all-reduced in-kernel over NVLink peer memory" if fused_allreduce
                                      else "ncclAllReduce") if world > 1 else
                                     "CUDA-graph replay")
                                  + "; one untimed throwaway solve of 36 iterations before the W warm-up steps "
                                    "(one-off initialisation: graph instantiation, peer mappings)"),
               "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
               "cpu_baseline": cpu, "resid_norm_after_timed_region": st.resid_norm}
        out.update(extra)
        print(json.dumps(out))
    if world > 1:
        ctx.barrier()


# ---------------------------------------------------------- reference arm
def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    # the workload is the one our arm runs at this N: under torchrun WORLD_SIZE says so, launched as a
    # plain process --gpus does
    world = int(os.environ.get("WORLD_SIZE", str(max(1, args.gpus))))
    if rank != 0:
        return
    g = args.grid or (G_CONFIG2 if world == 1 else G_CONFIG5)
    # The real operator of the configuration, in full, on the host (configs[4]: 10^8 rows = 6.4 GB of
    # CSR + 5 vectors; ~1.5 s per iteration): K iterations after W when that fits the time budget,
    # otherwise as many as fit (then `steps` says how many ran).  Nothing is extrapolated.
    try:
        res = run_reference_cg(g, max_steps=args.steps, warmup=args.warmup, budget_s=150.0)
        extrapolated = False
        value = res["value"]
        sample = res["sample"]
    except MemoryError:
        # host too small for the 10^8-row operator: sample the 10^7-row grid and scale by the row ratio
        # (a CG iteration is linear in the rows); flagged as such
        res = run_reference_cg(G_CONFIG2, max_steps=args.steps, warmup=args.warmup, budget_s=120.0)
        extrapolated = True
        value = res["value"] * (G_CONFIG2 * G_CONFIG2) / float(g * g)
        sample = res["sample"] + "; EXTRAPOLATED by rows %d/%d to the %dx%d workload" % (
            G_CONFIG2 * G_CONFIG2, g * g, g, g)
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
           "steps": res["steps"], "warmup": res["warmup"], "ms_per_step": 1e3 / value,
           "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": workload_config(g, world),
           "implementation": "reference pykrylov CG.solve on the host cores (CPU only, no GPU)",
           "extrapolated": extrapolated,
           "cpu_baseline": {"value": value, "unit": UNIT, "cores": res["cores"], "kind": res["kind"],
                            "sample": sample},
           "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=0, help="override the Laplacian grid size (debugging)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-single", action="store_true", help="skip the 1-GPU same-workload leg (N>1)")
    ap.add_argument("--no-config5", action="store_true",
                    help="N=1: skip the side leg that times the 10^8-row operator of configs[4] on this GPU")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="N=1: skip the side legs for BASELINE configs[2] (MINRES) and configs[3] (Bi-CGSTAB)")
    ap.add_argument("--cg-fuse", type=int, default=None, choices=[0, 1, 2],
                    help="CG launch plan (KRY_OPT_CG_FUSE); default: the library's default")
    ap.add_argument("--cg-fuse-shards", type=int, default=None, choices=[0, 1],
                    help="N>1: row shards use the fused CG plan too (KRY_OPT_CG_FUSE_SHARDS)")
    ap.add_argument("--halo-p2p", type=int, default=None, choices=[0, 1],
                    help="N>1: halo exchange through peer memory inside the SpMV launch (KRY_OPT_HALO_P2P)")
    ap.add_argument("--nccl-allreduce", action="store_true",
                    help="N>1: use ncclAllReduce + finalize launches instead of the fused NVLink peer-memory all-reduce")
    args = ap.parse_args()
    if args.impl == "reference":
        main_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        main_ours(args)


if __name__ == "__main__":
    main()
