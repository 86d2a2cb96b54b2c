#!/bin/bash
# round-2 job A: GPU tests + bench with the committed library, then the marching kernel (dev library) against round-1 kernels
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2a_tests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err
SYK_LIB_NAME=libsyk_dev.so timeout 300 python tools/cs_check.py --big --time > gpurun_out/r2a_cs_check.log 2>&1
tail -c 500 gpurun_out/r2a_tests.log; head -c 600 gpurun_out/bench_r2a.json; echo; tail -3 gpurun_out/bench_r2a.err; tail -30 gpurun_out/r2a_cs_check.log
