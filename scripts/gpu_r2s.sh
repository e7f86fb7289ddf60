#!/bin/bash
# Round 2, GPU call S (1 GPU): the whole -m gpu suite on the state with fused lls trips and the 16-byte
# multi-AXPY, then lls trip rates A/B (where the recurrence phase runs, fused vs three-launch form, graph vs enqueued).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2s_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2s_pytest_gpu.log | cut -c1-300
timeout 400 python scripts/gpu_lls_rates.py 2>&1 | tail -12 | tee gpurun_out/r2s_lls_rates.log
