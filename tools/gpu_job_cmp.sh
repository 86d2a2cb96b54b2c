#!/bin/bash
# usage: gpu_job_cmp.sh libA.so libB.so ...  -- bench step / k_cs_fast time per library, then cs parity tests with the first one
for v in "$@"; do
  SYK_LIB_NAME=$v python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-stress 2>/dev/null > gpurun_out/tmp_cmp.json
  python -c "import json; d=json.load(open('gpurun_out/tmp_cmp.json')); print('$v', d['ms_per_step'], d['roofline']['avg_launch_ms'], d['parity']['status'])"
done
SYK_LIB_NAME=$1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_exact.py -q -m gpu -k "cs or block or golden or known" 2>&1 | tail -2
