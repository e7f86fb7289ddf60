#!/bin/bash
# Round 2, GPU call I (1 GPU): final parity suite, bench line, launch list and ncu --set full captures.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r2i_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2i_pytest_gpu.log | cut -c1-300
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/r2i_clocks.csv &
SMI=$!
timeout 900 python bench.py > gpurun_out/r2i_bench_n1.json 2> gpurun_out/r2i_bench_n1.err; echo "bench rc=$?"
kill $SMI
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2i_bench_n1.json").read())
print("value %.1f e2e %.1f frac %.3f k1 %.4f ms cpu %.2f" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d["cpu_baseline"]["value"]))
for k in ("config3_minres", "config4_bicgstab", "config5_one_gpu"):
    print(k, json.dumps(d.get(k))[:600])
PY
# launch list of the bench command (per-launch times are cold-cache and serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2i_launches.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu --no-config5 --no-other-configs > gpurun_out/r2i_ncu_bench.log 2>&1
# --set full: the dominant kernel (fused CG SpMV launch) and the second CG launch
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_row_kernel -s 6 -c 2 -o gpurun_out/r2i_prof_cg_fused -f \
    python bench.py --steps 12 --warmup 3 --no-cpu --no-config5 --no-other-configs > gpurun_out/r2i_ncu_full_cg.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:vec_pass_kernel -s 6 -c 2 -o gpurun_out/r2i_prof_cg_pass -f \
    python bench.py --steps 12 --warmup 3 --no-cpu --no-config5 --no-other-configs > gpurun_out/r2i_ncu_full_cg2.log 2>&1
# --set full: one iteration of Bi-CGSTAB (config 4), MINRES (config 3), TFQMR (config 2 operator)
for s in bicgstab minres tfqmr; do
  timeout 600 ncu --set full --clock-control none -k regex:'spmv_row_kernel|vec_pass_kernel|vec_map_kernel' -s 8 -c 8 -o gpurun_out/r2i_prof_$s -f \
      python scripts/prof_solvers.py $s 6 > gpurun_out/r2i_ncu_full_$s.log 2>&1
done
for r in cg_fused cg_pass bicgstab minres tfqmr; do
  ncu -i gpurun_out/r2i_prof_$r.ncu-rep --page raw --csv > gpurun_out/r2i_prof_${r}_raw.csv 2>/dev/null
done
ncu -i gpurun_out/r2i_prof_cg_fused.ncu-rep --page source --csv > gpurun_out/r2i_prof_cg_fused_source.csv 2>/dev/null
# the reports themselves are too large to travel back (64 MiB cap): the CSV pages are what is kept
rm -f gpurun_out/r2i_prof_*.ncu-rep
cat gpurun_out/r2i_ncu_full_minres.log | tail -5
ls -la gpurun_out | head -40
