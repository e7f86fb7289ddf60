#!/bin/bash
# Round 2, GPU call L (1 GPU): compute-sanitizer memcheck over the new kernels (assembly, algebra, lls step
# kernels, one-CTA CG, MINRES plans) on the small test cases.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file gpurun_out/r2l_memcheck.log \
    python -m pytest tests/test_gpu_lls.py tests/test_gpu_parity.py -m gpu -q --timeout 1200 \
    -k "lsqr or lsmr or craig or symmlq or coord_operator or operator_algebra or one_cta or minres_single_step or spmv_bit_exact or cg_single_step or closure" \
    > gpurun_out/r2l_pytest_memcheck.log 2>&1; echo "memcheck pytest rc=$?"
tail -4 gpurun_out/r2l_pytest_memcheck.log | cut -c1-300
grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/r2l_memcheck.log | sort | uniq -c | head
