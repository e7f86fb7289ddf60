#!/bin/bash
# Round 2, GPU call O (8 GPUs): final state -- sharded parity test at 2/4/8 ranks, bench with the fused exchange,
# and the ranks' own SpMV launch times with every exchange taken out of the kernel (skew).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 1200 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 1000 -rs > gpurun_out/r2o_pytest_multi_gpu_n$N.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2o_pytest_multi_gpu_n$N.log
grep -h "world\|ok" gpurun_out/multi_gpu_worker_n$N.log | tail -20
run_bench () {
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
      bench.py --gpus $N --steps 100 --warmup 10 "$@" > gpurun_out/r2o_bench_n${N}_$tag.json 2> gpurun_out/r2o_bench_n${N}_$tag.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2o_bench_n${N}_$tag.json") if l.startswith("{")][-1])
    r = d["roofline"]
    print("$tag: value %.1f it/s  ms/step %.4f  k1 %.4f ms frac %.3f  e2e %.1f  speedup %.2f  resid %r one-gpu %r" % (
        d["value"], d["ms_per_step"], r["avg_launch_ms"], r["frac"], d["e2e"]["value"], d.get("speedup_vs_one_gpu", 0),
        d["resid_norm_after_timed_region"], d.get("one_gpu_same_workload", {}).get("resid_norm_after_timed_region")))
    print("   per rank k1 ms:", ["%.4f" % v for v in r.get("avg_launch_ms_per_rank", [])])
except Exception as e:
    print("$tag: no line:", e)
PY
}
run_bench halo1
run_bench halo0 --halo-p2p 0 --no-single
