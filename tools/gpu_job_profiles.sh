#!/bin/bash
# round-2 profiles: launch list, full captures of k_cs_fast and k_scan inside the bench
B="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-stress"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r2.csv $B > gpurun_out/bench_under_ncu_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_cs_fast -s 16 -c 1 -f -o gpurun_out/csfast_r2 $B > gpurun_out/ncu_cs_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scan -s 24 -c 3 -f -o gpurun_out/scan_r2 $B > gpurun_out/ncu_scan_r2.log 2>&1
cp syconn_b200/libsyk.so gpurun_out/libsyk_r2.so
tail -2 gpurun_out/ncu_cs_r2.log gpurun_out/ncu_scan_r2.log; wc -l gpurun_out/launches_r2.csv
