#!/bin/bash
python tools/ccl_debug.py > gpurun_out/r2d_ccl.log 2>&1
for v in a b c d; do SYK_LIB_NAME=libsyk_sp$v.so python tools/org_time.py >> gpurun_out/r2d_times.log 2>&1; done
cat gpurun_out/r2d_ccl.log gpurun_out/r2d_times.log
