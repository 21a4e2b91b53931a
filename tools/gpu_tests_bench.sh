#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== host overhead"; python tools/host_overhead.py 2>&1 | tail -2
echo "== bench ours"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_ours.json; python -c "
import json
d=json.load(open('gpurun_out/bench_ours.json'))
print('fwd_ms',d['fwd_ms'],'step_ms',d['ms_per_step'],'e2e',d['e2e']['value'], 1000/d['e2e']['value'])
print({k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})"
