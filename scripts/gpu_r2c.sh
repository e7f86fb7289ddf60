#!/bin/bash
# Round 2, GPU call C (N GPUs): the sharded path's hardware parity test with the shipped defaults, log kept.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 1500 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 1200 -rs > gpurun_out/r2c_pytest_multi_gpu_n$N.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r2c_pytest_multi_gpu_n$N.log
for f in gpurun_out/multi_gpu_worker_n*.log; do echo "== $f"; grep -v "^$" $f | tail -30; done
