#!/bin/bash
# Round 2, GPU call X (1 GPU): compute-sanitizer memcheck over the kernels added last (SpMV with the fused
# y-side update, 16-byte multi-AXPY body incl. odd lengths, fused lls trips and their graph replay).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 55 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file gpurun_out/r2x_memcheck.log \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_lls.py -m gpu -q --timeout 50 \
    -k "multi_axpy or fused_y_side or fused_trip or trip_graph" \
    > gpurun_out/r2x_pytest_memcheck.log 2>&1; echo "memcheck pytest rc=$?"
tail -3 gpurun_out/r2x_pytest_memcheck.log | cut -c1-300
grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/r2x_memcheck.log | sort | uniq -c | head
