#!/bin/bash
# Final GPU pass of round 1 (ONE GPU, every step bounded).  Outputs land in gpurun_out/.
mkdir -p gpurun_out
echo "== pytest gpu"; date
timeout 300 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
echo "== bench"; date
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
timeout 240 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
kill $SMI
cut -c1-300 gpurun_out/bench_n1.json
echo "== A/B"; date
timeout 120 python scripts/gpu_ab_cgfuse.py 3162 > gpurun_out/ab_cgfuse.log 2>&1; echo "ab rc=$?"; grep hints1 gpurun_out/ab_cgfuse.log
echo "== ncu launch list"; date
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu list rc=$?"
echo "== ncu full"; date
timeout 240 ncu --set full --clock-control none --import-source on -k regex:'spmv_row_kernel|vec_pass_kernel' -s 8 -c 4 \
    -o gpurun_out/prof_cg_fused -f python bench.py --steps 12 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/prof_cg_fused.ncu-rep --page raw --csv > gpurun_out/prof_cg_fused_raw.csv 2>/dev/null
echo "== configs"; date
timeout 240 python scripts/bench_configs.py > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; echo "configs rc=$?"
cut -c1-400 gpurun_out/configs.jsonl; tail -5 gpurun_out/configs.err
date
