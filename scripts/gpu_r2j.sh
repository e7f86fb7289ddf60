#!/bin/bash
# Round 2, GPU call J (1 GPU): final parity suite (+ measured per-step margins), solver rates.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2j_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2j_pytest_gpu.log | cut -c1-300
timeout 600 python scripts/gpu_solver_rates.py 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()"
