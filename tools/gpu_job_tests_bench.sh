#!/bin/bash
python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2k_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2k.json 2> gpurun_out/bench_r2k.err
tail -c 500 gpurun_out/r2k_tests.log; head -c 300 gpurun_out/bench_r2k.json; echo; tail -5 gpurun_out/bench_r2k.err
