#!/bin/bash
export SYK_LIB_NAME=libsyk_dev.so
python tools/cs_time.py > gpurun_out/r2j_times.log 2>&1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_exact.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2j_tests.log
SYK_CS_DEBUG=1 python tools/cs_once.py >> gpurun_out/r2j_times.log 2>&1
cat gpurun_out/r2j_times.log | tail -8; tail -c 800 gpurun_out/r2j_tests.log
