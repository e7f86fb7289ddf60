#!/bin/bash
# Round 2, GPU call R (1 GPU): final parity suite of the committed state, lls trip rates (graph vs enqueued),
# smoke, a bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2r_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2r_pytest_gpu.log | cut -c1-300
timeout 200 python scripts/gpu_lls_rates.py 2>&1 | tail -12 | tee gpurun_out/r2r_lls_rates.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 240 python bench.py --no-other-configs > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"; cut -c1-1200 gpurun_out/r2r_bench.json
