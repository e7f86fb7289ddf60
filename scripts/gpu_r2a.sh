#!/bin/bash
# Round 2, GPU call A (1 GPU): the merged candidates -- parity suite, kernel sweep, configs 0/2/3, bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2a_pytest_gpu.log
timeout 600 python scripts/gpu_tune.py > gpurun_out/r2a_tune.log 2>&1; tail -30 gpurun_out/r2a_tune.log
timeout 600 python scripts/bench_configs.py 0 2 3 > gpurun_out/r2a_configs.jsonl 2> gpurun_out/r2a_configs.err; cut -c1-900 gpurun_out/r2a_configs.jsonl; tail -5 gpurun_out/r2a_configs.err
timeout 600 python bench.py > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; cut -c1-1500 gpurun_out/r2a_bench_n1.json
