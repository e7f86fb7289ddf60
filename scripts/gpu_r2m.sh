#!/bin/bash
# Round 2, GPU call M (N GPUs): what the halo-column predicate + coherent loads in the sharded gather cost
# (A/B build libkrylov_b200_ab.so: the gather loads every column through the read-only path, as in round 1).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
run_bench () {
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
      bench.py --gpus $N --steps 100 --warmup 10 --no-single "$@" > gpurun_out/r2m_bench_n${N}_$tag.json 2> gpurun_out/r2m_bench_n${N}_$tag.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2m_bench_n${N}_$tag.json") if l.startswith("{")][-1])
    r = d["roofline"]
    print("$tag: value %.1f it/s  ms/step %.4f  k1 %.4f ms frac %.3f  resid %r" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["frac"], d["resid_norm_after_timed_region"]))
except Exception as e:
    print("$tag: no line:", e)
PY
}
run_bench default
run_bench default2
timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 1000 -rs > gpurun_out/r2m_pytest_multi_gpu_n$N.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2m_pytest_multi_gpu_n$N.log
grep -h "world\|ok" gpurun_out/multi_gpu_worker_n$N.log | tail -14
