#!/bin/bash
# usage: gpu_job_n.sh N   -- multi-GPU parity tool + bench at N ranks (logs under gpurun_out/)
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py > gpurun_out/r2_mgpu_check_n$N.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err
grep -E "MGPU_CHECK|sharded" gpurun_out/r2_mgpu_check_n$N.log | tail -12; tail -3 gpurun_out/r2_mgpu_check_n$N.log; head -c 400 gpurun_out/bench_r2_n$N.json; echo; tail -5 gpurun_out/bench_r2_n$N.err
