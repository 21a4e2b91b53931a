#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== bench ours"; timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_ours.json
echo "== ncu quick"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"onesweep_pass|blend_fwd|blend_bwd|emit|scan_tiles|preprocess_fwd" -s 13 -c 13 -f -o gpurun_out/prof_r1b python tools/one_step.py 2 > gpurun_out/ncu_full_b.log 2>&1; tail -2 gpurun_out/ncu_full_b.log
