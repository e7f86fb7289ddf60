#!/bin/bash
# Round 2, GPU call Q (2 GPUs): ncu --set full of the SHARD kernels.  The two ranks are started by hand, rank 0
# under ncu; every exchange goes through NCCL (--halo-p2p 0 --nccl-allreduce), so the profiled kernel has no
# peer dependency and can be replayed while rank 1 waits in the next collective.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export MASTER_ADDR=127.0.0.1 MASTER_PORT=29611 WORLD_SIZE=2 KRY_RENDEZVOUS_TAG=r2q_$$
ARGS="bench.py --gpus 2 --steps 6 --warmup 3 --no-single --no-cpu --halo-p2p 0 --nccl-allreduce"
RANK=1 LOCAL_RANK=1 timeout 900 python $ARGS > gpurun_out/r2q_rank1.log 2>&1 &
P1=$!
RANK=0 LOCAL_RANK=0 timeout 900 ncu --set full --clock-control none -k regex:'spmv_row_shard_kernel|vec_pass_kernel' -s 8 -c 4 \
    -o gpurun_out/r2q_prof_shard -f python $ARGS > gpurun_out/r2q_rank0_ncu.log 2>&1
echo "rank0 rc=$?"
wait $P1; echo "rank1 rc=$?"
ncu -i gpurun_out/r2q_prof_shard.ncu-rep --page raw --csv > gpurun_out/r2q_prof_shard_raw.csv 2>/dev/null
rm -f gpurun_out/r2q_prof_shard.ncu-rep
tail -3 gpurun_out/r2q_rank0_ncu.log | cut -c1-300; ls -la gpurun_out/r2q_*
