#!/bin/bash
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2f_tests.log
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err
tail -c 1200 gpurun_out/r2f_tests.log; head -c 300 gpurun_out/bench_r2f.json; echo; tail -3 gpurun_out/bench_r2f.err
