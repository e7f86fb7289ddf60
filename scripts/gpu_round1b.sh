#!/bin/bash
# Second-session GPU pass (run under gpurun on ONE GPU): full GPU test-suite, CG launch-plan A/B,
# the secondary configs, the headline bench.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== pytest gpu"; date
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
echo "== smoke"; date
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== A/B cg fuse"; date
timeout 600 python scripts/gpu_ab_cgfuse.py > gpurun_out/ab_cgfuse.log 2>&1; echo "ab rc=$?"; cat gpurun_out/ab_cgfuse.log | tail -20
echo "== bench"; date
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-900 gpurun_out/bench_n1.json
for f in 1 2; do
  timeout 300 python bench.py --cg-fuse $f --no-cpu > gpurun_out/bench_n1_fuse$f.json 2> gpurun_out/bench_n1_fuse$f.err; echo "bench fuse$f rc=$?"; cut -c1-400 gpurun_out/bench_n1_fuse$f.json
done
echo "== configs"; date
timeout 900 python scripts/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?"; cut -c1-600 gpurun_out/configs.jsonl
date
