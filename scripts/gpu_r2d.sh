#!/bin/bash
# Round 2, GPU call D (N GPUs): sharded parity test + bench A/B of the halo paths on the 10^8-row operator.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
STEPS=${STEPS:-100}
timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 1000 -rs > gpurun_out/r2d_pytest_multi_gpu_n$N.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2d_pytest_multi_gpu_n$N.log
for f in gpurun_out/multi_gpu_worker_n*.log; do echo "== $f"; grep -v "^$" $f | tail -16; done
run_bench () {  # $1 = tag, rest = flags
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
      bench.py --gpus $N --steps $STEPS --warmup 10 "$@" > gpurun_out/r2d_bench_n${N}_$tag.json 2> gpurun_out/r2d_bench_n${N}_$tag.err
  echo "== bench $tag rc=$?"; python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2d_bench_n${N}_$tag.json") if l.startswith("{")][-1])
    r = d["roofline"]
    print("value %.1f it/s  ms/step %.4f  k1 %.4f ms frac %.3f  e2e %.1f  speedup %.2f  resid %r one-gpu %r" % (
        d["value"], d["ms_per_step"], r["avg_launch_ms"], r["frac"], d["e2e"]["value"], d.get("speedup_vs_one_gpu", 0),
        d["resid_norm_after_timed_region"], d.get("one_gpu_same_workload", {}).get("resid_norm_after_timed_region")))
except Exception as e:
    print("no line:", e)
PY
  tail -3 gpurun_out/r2d_bench_n${N}_$tag.err
}
run_bench halo1
run_bench halo0 --halo-p2p 0 --no-single
[ "$N" -lt 8 ] && run_bench halo1_again --no-single
run_bench nccl_allreduce --halo-p2p 0 --nccl-allreduce --no-single
KRY_HALO_TRACE=1 run_bench halo1_trace --no-single; grep "halo trace" gpurun_out/r2d_bench_n${N}_halo1_trace.err
