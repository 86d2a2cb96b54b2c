#!/bin/bash
# round-2 job C: sparse organelle kernel, CCL (f4), everything else
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2c_tests.log
python tools/org_time.py > gpurun_out/r2c_times.log 2>&1
SYK_ORG_SCAN=1 python tools/org_time.py >> gpurun_out/r2c_times.log 2>&1
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err
tail -c 700 gpurun_out/r2c_tests.log; cat gpurun_out/r2c_times.log; head -c 300 gpurun_out/bench_r2c.json; echo; tail -3 gpurun_out/bench_r2c.err
