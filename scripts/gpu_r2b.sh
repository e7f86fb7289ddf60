#!/bin/bash
# Round 2, GPU call B (1 GPU): full parity suite incl. the full-size config 3/4 tests; MINRES plan A/B.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2b_pytest_gpu.log
timeout 600 python scripts/gpu_minres_ab.py > gpurun_out/r2b_minres_ab.log 2>&1; tail -12 gpurun_out/r2b_minres_ab.log
