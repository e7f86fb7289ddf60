#!/bin/bash
# Round-1 profiling recipe (run under gpurun on ONE GPU). Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks.csv &
SMI=$!
python bench.py --steps 200 --warmup 10 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
kill $SMI
# every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
# full sets of the three CG kernels
ncu --set full --clock-control none --import-source on -k regex:spmv_row_kernel -s 4 -c 2 \
    -o gpurun_out/prof_spmv_dot -f python bench.py --steps 12 --warmup 3 --no-cpu > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vec_pass_kernel -s 4 -c 2 \
    -o gpurun_out/prof_cg_update -f python bench.py --steps 12 --warmup 3 --no-cpu > gpurun_out/ncu_full2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:vec_map_kernel -s 4 -c 2 \
    -o gpurun_out/prof_cg_dir -f python bench.py --steps 12 --warmup 3 --no-cpu > gpurun_out/ncu_full3.log 2>&1
ls -la gpurun_out
