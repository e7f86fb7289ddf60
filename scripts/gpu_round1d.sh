#!/bin/bash
# Last GPU pass of round 1: run with `gpurun --gpus 2`.  Sharded path after the solver / context
# changes of this session, the new two-region bench on one GPU, smoke.  Every step bounded.
mkdir -p gpurun_out
echo "== smoke"; date
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== bench n1"; date
timeout 200 python bench.py --no-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 rc=$?"; cut -c1-250 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
echo "== two-GPU test"; date
timeout 200 python -m pytest tests/test_multi_gpu.py -m gpu -q --timeout 150 > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest 2gpu rc=$?"; tail -3 gpurun_out/pytest_2gpu.log
echo "== bench n2"; date
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 60 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
cut -c1-250 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
date
