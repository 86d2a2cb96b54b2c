#!/bin/bash
# round-2 job B: TMA plane loads in k_cs_fast, empty-batch fast path in the organelle scan, aligned contact buffer
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2b_tests.log
python tools/cs_time.py > gpurun_out/r2b_times.log 2>&1
SYK_CS_NO_TMA=1 python tools/cs_time.py >> gpurun_out/r2b_times.log 2>&1
python tools/org_time.py >> gpurun_out/r2b_times.log 2>&1
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err
tail -c 400 gpurun_out/r2b_tests.log; cat gpurun_out/r2b_times.log; head -c 300 gpurun_out/bench_r2b.json; echo; tail -3 gpurun_out/bench_r2b.err
