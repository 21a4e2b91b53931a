#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_surface_gpu.py -m gpu -q --tb=short 2>&1 | tail -12
for mb in 6 7 8; do
  GRPG_BWD_MINB=$mb timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_mb$mb.json
  python -c "
import json
d=json.load(open('gpurun_out/bench_mb$mb.json'))
print('MINB=$mb step_ms',d['ms_per_step'],'e2e',d['e2e']['value'],'bwd',d['kernels']['blend_bwd']['ms_per_step'])"
done
python tools/gpu_check.py 2>&1 | tail -3 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['label'], {k:('%.2e'%v) for k,v in d.items() if k.startswith('rel_')})"
