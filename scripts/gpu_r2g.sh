#!/bin/bash
# Round 2, GPU call G (1 GPU): full parity suite; configs 0/2.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2g_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2g_pytest_gpu.log | cut -c1-300
timeout 600 python scripts/bench_configs.py 0 2 > gpurun_out/r2g_configs.jsonl 2> gpurun_out/r2g_configs.err; cut -c1-800 gpurun_out/r2g_configs.jsonl
