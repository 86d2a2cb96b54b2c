#!/bin/bash
python -m pytest tests/test_dropin.py tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -6 > gpurun_out/r2l_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r2l_n2.json 2> gpurun_out/bench_r2l_n2.err
tail -c 400 gpurun_out/r2l_tests.log; head -c 200 gpurun_out/bench_r2l_n2.json; echo; tail -3 gpurun_out/bench_r2l_n2.err
