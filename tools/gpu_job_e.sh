#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2e_tests.log
python tools/cs_time.py > gpurun_out/r2e_times.log 2>&1
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err
tail -c 900 gpurun_out/r2e_tests.log; cat gpurun_out/r2e_times.log; head -c 300 gpurun_out/bench_r2e.json; echo; tail -3 gpurun_out/bench_r2e.err
