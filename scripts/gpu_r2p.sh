#!/bin/bash
# Round 2, GPU call P (N GPUs): who publishes the boundary entries -- the last row CTAs (0) or dedicated CTAs
# that own no rows (KRY_HALO_PUSH_CTAS = 8 / 2); same box, alternating runs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
run_bench () {
  tag=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) \
      bench.py --gpus $N --steps 100 --warmup 10 --no-single "$@" > gpurun_out/r2p_bench_n${N}_$tag.json 2> gpurun_out/r2p_bench_n${N}_$tag.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r2p_bench_n${N}_$tag.json") if l.startswith("{")][-1])
    r = d["roofline"]
    print("$tag: value %.1f it/s  ms/step %.4f  k1 %.4f ms frac %.3f  resid %r" % (d["value"], d["ms_per_step"], r["avg_launch_ms"], r["frac"], d["resid_norm_after_timed_region"]))
except Exception as e:
    print("$tag: no line:", e)
PY
}
KRY_HALO_PUSH_CTAS=0 run_bench push_by_row_ctas
KRY_HALO_PUSH_CTAS=8 run_bench dedicated8
KRY_HALO_PUSH_CTAS=2 run_bench dedicated2
if [ "$N" -lt 8 ]; then
KRY_HALO_PUSH_CTAS=0 run_bench push_by_row_ctas_again
KRY_HALO_PUSH_CTAS=8 run_bench dedicated8_again
fi
