#!/bin/bash
for v in libsyk_dev.so libsyk_lut.so; do
  SYK_LIB_NAME=$v python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-stress 2>/dev/null > gpurun_out/tmp_lu.json
  python -c "import json; d=json.load(open('gpurun_out/tmp_lu.json')); print('$v', d['ms_per_step'], d['roofline']['avg_launch_ms'], d['parity']['status'])"
done
SYK_LIB_NAME=libsyk_dev.so python -m pytest tests/test_gpu_parity.py -q -m gpu -k "cs or block or golden or known" 2>&1 | tail -2
