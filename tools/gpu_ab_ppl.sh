#!/bin/bash
mkdir -p gpurun_out
for v in 8 10 12; do
GRPG_FWD_MINB=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_f$v.json
python -c "
import json
d=json.load(open('gpurun_out/bench_f$v.json'))
print('FWD_MINB=$v fwd_ms',d['fwd_ms'],'step_ms',d['ms_per_step'],'blend_fwd',d['kernels']['blend_fwd']['ms_per_step'])"
done
