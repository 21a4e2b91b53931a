#!/bin/bash
mkdir -p gpurun_out
for mb in 1 3 4 5; do
  GRPG_PBWD_MINB=$mb timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_pb$mb.json
  python -c "
import json
d=json.load(open('gpurun_out/bench_pb$mb.json'))
print('MINB=$mb step_ms',d['ms_per_step'],'pbwd',d['kernels']['preprocess_bwd']['ms_per_step'])"
done
