#!/bin/bash
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2g_tests.log
tail -c 1500 gpurun_out/r2g_tests.log
