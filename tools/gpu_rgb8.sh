#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_image_gpu.py -m gpu -q 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_ours.json; python -c "
import json
d=json.load(open('gpurun_out/bench_ours.json'))
print('fwd_ms',d['fwd_ms'],'step_ms',d['ms_per_step'],'e2e',d['e2e']['value'],'e2e_fwd',d['e2e_fwd']['value'],'rgb8',d['e2e_fwd_rgb8']['value'])"
