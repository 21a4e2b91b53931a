#!/bin/bash
# A/B of the wide backward blend's occupancy targets (PPL=2): parity + bench kernel table for each.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q --tb=line -x -k "golden or oracle" 2>&1 | tail -2
for mb in 5 6 7 8; do
  GRPG_BWD_MINB=$mb timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_mb$mb.json
  python -c "
import json
d=json.load(open('gpurun_out/bench_mb$mb.json'))
print('MINB=$mb step_ms',d['ms_per_step'],'e2e',d['e2e']['value'],'bwd',d['kernels']['blend_bwd']['ms_per_step'])"
done
