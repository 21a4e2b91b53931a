#!/bin/bash
# Trimmed refresh after a kernel change: GPU tests, our bench line, ncu launch list and the ncu --set full capture.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "== bench ours"; timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_ours.json; python -c "
import json
d=json.load(open('gpurun_out/bench_ours.json'))
print('fwd_ms',d['fwd_ms'],'step_ms',d['ms_per_step'],'e2e',d['e2e']['value'])
print({k:round(v['ms_per_step'],4) for k,v in d['kernels'].items()})"
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches.log 2>&1; tail -1 gpurun_out/launches.log | cut -c1-120
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blend_fwd|blend_bwd|onesweep_pass|radix_tile|emit_instances|preprocess_fwd|preprocess_bwd|scan_tiles|tile_ranges" -s 19 -c 19 -f -o gpurun_out/prof_final python tools/one_step.py 2 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
