#!/bin/bash
# Round 2, GPU call K (1 GPU): TMA-staged row kernel -- bit-exactness and A/B against the row kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "spmv or fullsize_kernels" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2k_pytest.log | cut -c1-300
timeout 600 python scripts/gpu_tune.py > gpurun_out/r2k_tune.log 2>&1; grep -E "^row/|^rowpf" gpurun_out/r2k_tune.log
