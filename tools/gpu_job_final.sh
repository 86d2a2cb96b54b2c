#!/bin/bash
# final verification of a round: smoke, the GPU tests, the bench (both arms), launch list of the worker body
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -1
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; head -c 200 gpurun_out/bench_final.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; head -c 300 gpurun_out/bench_final_ref.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/worker_launches.csv python tools/worker_time.py 1 > /dev/null 2>&1
