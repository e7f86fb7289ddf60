#!/bin/bash
# Round 2, GPU call E (2 GPUs): where does the fused-halo SpMV launch lose time? (in-kernel timers)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
run_bench () {
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
      bench.py --gpus $N --steps 60 --warmup 10 --no-single "$@" > gpurun_out/r2e_bench_n${N}_$tag.json 2> gpurun_out/r2e_bench_n${N}_$tag.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2e_bench_n${N}_$tag.json") if l.startswith("{")][-1])
    r = d["roofline"]
    print("$tag: value %.1f it/s  ms/step %.4f  k1 %.4f ms frac %.3f  resid %r" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["frac"], d["resid_norm_after_timed_region"]))
except Exception as e:
    print("$tag: no line:", e)
PY
  grep "halo trace" gpurun_out/r2e_bench_n${N}_$tag.err
}
KRY_HALO_TRACE=1 run_bench halo1_trace
