#!/bin/bash
# Round 2, GPU call S (1 GPU): lls / SYMMLQ trips with the y-side update, inner product and recurrence
# phase fused into the SpMV launch: the lls GPU tests + the new SpMV-epilogue test, then trip rates A/B.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lls.py tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "lls or lsqr or lsmr or craig or symmlq or fused or trip or closure or multi_axpy" > gpurun_out/r2s_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2s_pytest_gpu.log | cut -c1-300
timeout 300 python scripts/gpu_lls_rates.py 2>&1 | tail -12 | tee gpurun_out/r2s_lls_rates.log
