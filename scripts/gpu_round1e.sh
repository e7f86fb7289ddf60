#!/bin/bash
# 2-GPU pass: fused CG plan on row shards (KRY_OPT_CG_FUSE_SHARDS) -- parity test, then A/B.
mkdir -p gpurun_out
echo "== two-GPU test"; date
timeout 200 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 150 > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest 2gpu rc=$?"; tail -15 gpurun_out/pytest_2gpu.log
for fs in 0 1; do
  echo "== bench n2 fuse_shards=$fs"; date
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$fs \
      bench.py --gpus 2 --steps 60 --warmup 5 --no-single --cg-fuse-shards $fs > gpurun_out/bench_n2_fs$fs.json 2> gpurun_out/bench_n2_fs$fs.err; echo "rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n2_fs$fs.json').read().splitlines()[-1])
print('fuse_shards=$fs value', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'plan', d['roofline']['cg_launch_plan'], 'k1 ms', round(d['roofline']['avg_launch_ms'],4), 'frac', round(d['roofline']['frac'],3))
PY
done
date
