#!/bin/bash
# Round 2, GPU call F (1 GPU): full parity suite, MINRES plan A/B (timing fixed), configs 0/2, bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2f_pytest_gpu.log
timeout 600 python scripts/gpu_minres_ab.py > gpurun_out/r2f_minres_ab.log 2>&1; tail -6 gpurun_out/r2f_minres_ab.log | cut -c1-420
timeout 600 python scripts/bench_configs.py 0 2 > gpurun_out/r2f_configs.jsonl 2> gpurun_out/r2f_configs.err; cut -c1-700 gpurun_out/r2f_configs.jsonl
timeout 600 python bench.py --no-config5 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; cut -c1-300 gpurun_out/r2f_bench_n1.json; python -c "
import json; d=json.loads(open('gpurun_out/r2f_bench_n1.json').read()); print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'],'cpu',d['cpu_baseline']['value'])"
