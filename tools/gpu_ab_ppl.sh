#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q --tb=short 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_ours.json
python -c "
import json
d=json.load(open('gpurun_out/bench_ours.json'))
print('step_ms',d['ms_per_step'],'e2e',d['e2e']['value'],'bwd',d['kernels']['blend_bwd']['ms_per_step'])"
