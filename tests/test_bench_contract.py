"""bench.py contract, as far as it can be checked without a GPU: the reference arm runs the
reference's CPU path and prints exactly one JSON line with the agreed keys; the other ranks of
a torchrun launch stay silent; our own arm refuses to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def run(args, env=None, timeout=300):
    e = dict(os.environ)
    e.pop("RANK", None)
    e.pop("WORLD_SIZE", None)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, env=e, timeout=timeout)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = run(["--impl", "reference", "--grid", "64", "--steps", "4", "--warmup", "1"])
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "cg_iters_per_s" and d["unit"] == "iters/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["dtype"] == "f64"
    assert d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["value"] > 0 and abs(d["ms_per_step"] - 1e3 / d["value"]) <= 1e-9 * d["ms_per_step"]
    assert d["config"]["rows"] == 64 * 64 and d["config"]["nnz"] == 5 * 64 * 64 - 4 * 64
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_silently():
    r = run(["--impl", "reference", "--gpus", "2", "--grid", "64", "--steps", "2"],
            env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_fails_loudly_without_a_gpu():
    import pykrylov_b200.device as dev
    try:
        n = dev.device_count()
    except Exception:
        n = 0
    if n > 0:
        import pytest
        pytest.skip("a GPU is present")
    r = run(["--steps", "2", "--grid", "32", "--no-cpu"], timeout=120)
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "libkrylov_b200 error" in r.stderr


def test_algorithmic_byte_model():
    sys.path.insert(0, ROOT)
    import bench
    g = 3162
    n, nnz = g * g, 5 * g * g - 4 * g
    assert bench.spmv_bytes(n, nnz) == 799707748
    assert bench.cg_iter_bytes(n, nnz) == 1519581316
    assert bench.spmv_bytes(n, nnz) + bench.K1_EXTRA_BYTES_PER_ROW[2] * n == 1119651556
    for form in (0, 1, 2):                                   # what the two/three launches of a plan move
        assert bench.ITER_VECTOR_BYTES_PER_ROW[form] == 72 - 8 * form


def test_our_arm_code_path_under_emulation(emu_ctx, capsys, monkeypatch):
    """bench.main_ours end to end -- both timed regions, the e2e solve through the public API, the
    config-5 side leg, the JSON line -- on the host emulation of the device logic with toy grids.
    The numbers mean nothing here; what is checked is that every leg runs and the line is complete."""
    import argparse
    sys.path.insert(0, ROOT)
    import bench
    import pykrylov_b200.comm as comm
    monkeypatch.setattr(bench, "G_CONFIG2", 24)
    monkeypatch.setattr(bench, "G_CONFIG5", 40)
    monkeypatch.setattr(comm, "init_from_env", lambda *a, **k: (emu_ctx, 0, 1))
    args = argparse.Namespace(gpus=1, steps=30, warmup=3, impl="ours", grid=0, no_cpu=True, no_single=False,
                              no_config5=False, no_other_configs=True, cg_fuse=None, cg_fuse_shards=None, nccl_allreduce=False, halo_p2p=None)
    bench.main_ours(args)
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches",
                "roofline", "cpu_baseline", "config5_one_gpu"):
        assert key in d, key
    assert d["metric"] == "cg_iters_per_s" and d["steps"] == 30 and d["n_gpus"] == 1 and d["value"] > 0
    assert d["config"]["rows"] == 24 * 24 and d["config"]["nnz"] == 5 * 24 * 24 - 4 * 24
    assert d["gpu_launches"] == 2 * 30                               # the 2-launch CG plan
    assert d["roofline"]["cg_launch_plan"] == 2 and d["roofline"]["bound"] == "hbm"
    assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    side = d["config5_one_gpu"]
    assert "error" not in side and side["value"] > 0 and side["rows"] == 40 * 40


def test_smoke_logic_under_emulation(emu_ctx, capsys):
    """__graft_entry__.smoke() -- the driver's first GPU step -- on the emulated device: its own
    assertions (bit-exact SpMV against the oracle, history and solution within tolerance, the
    launch count of the 2-launch CG plan) are part of what is checked."""
    sys.path.insert(0, ROOT)
    import __graft_entry__ as entry
    entry.smoke()
    assert "smoke ok" in capsys.readouterr().out


def test_our_arm_multi_rank_code_path_under_emulation(emu_ctx, capsys, monkeypatch):
    """bench.main_ours as the driver launches it for N > 1 (one rank per GPU, row-sharded operator,
    max-over-ranks timing, the 1-GPU same-workload leg on rank 0), with three emulated ranks as
    threads of this process.  Numbers are meaningless here; every leg must run, only rank 0 may
    print, and the line must say which launch plan the shards ran."""
    import argparse
    import ctypes as C
    import threading
    sys.path.insert(0, ROOT)
    import bench
    import pykrylov_b200.comm as comm
    from pykrylov_b200 import _lib as L
    from pykrylov_b200 import device as dev
    world = 3
    monkeypatch.setattr(bench, "G_CONFIG5", 30)
    uid = C.create_string_buffer(L.KRY_COMM_ID_BYTES)
    L.call("kry_comm_unique_id", uid)
    local = threading.local()

    def init_from_env(*a, **k):
        ctx = dev.Context(0)
        ctx.comm_init(world, local.rank, uid.raw)
        local.ctx = ctx
        return ctx, local.rank, world

    monkeypatch.setattr(comm, "init_from_env", init_from_env)
    errors = [None] * world

    def body(rank):
        local.rank = rank
        args = argparse.Namespace(gpus=world, steps=24, warmup=3, impl="ours", grid=0, no_cpu=True, no_single=False,
                                  no_config5=False, no_other_configs=True, cg_fuse=None, cg_fuse_shards=None, nccl_allreduce=False, halo_p2p=None)
        try:
            bench.main_ours(args)
        except BaseException as exc:                    # noqa: BLE001
            errors[rank] = exc
        finally:
            if getattr(local, "ctx", None) is not None:
                local.ctx.close()

    threads = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(240)
    assert not any(t.is_alive() for t in threads), "a rank is stuck"
    assert errors == [None] * world, errors
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.strip()]
    assert len(lines) == 1                                           # rank 0 only
    d = json.loads(lines[0])
    assert d["n_gpus"] == world and d["scaling"] == "strong" and d["value"] > 0
    assert d["config"]["rows"] == 30 * 30 and d["config"]["rows_per_gpu"] == 300
    assert d["roofline"]["cg_launch_plan"] == 2 and d["gpu_launches"] >= 2 * 24      # fused plan on shards (+ halo packs)
    assert d["one_gpu_same_workload"]["value"] > 0 and d["speedup_vs_one_gpu"] > 0
    assert d["e2e"]["value"] > 0 and "config5_one_gpu" not in d
