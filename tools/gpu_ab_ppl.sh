#!/bin/bash
# A/B of the wide backward blend's occupancy targets (PPL=2): bench kernel table for each.
mkdir -p gpurun_out
for mb in 4 5 6 7 8; do
  echo "== MINB=$mb"
  GRPG_BWD_PPL=2 GRPG_BWD_MINB=$mb timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_mb$mb.json
  python -c "
import json
d=json.load(open('gpurun_out/bench_mb$mb.json'))
print('fwd_ms',d['fwd_ms'],'step_ms',d['ms_per_step'],'e2e',d['e2e']['value'],'bwd',d['kernels']['blend_bwd'])"
done
