#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python tools/one_step.py 3 > gpurun_out/launches.log 2>&1; tail -2 gpurun_out/launches.log
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blend_fwd|blend_bwd|onesweep_pass|emit_instances|preprocess_fwd|preprocess_bwd|scan_tiles|tile_ranges" -s 16 -c 16 -f -o gpurun_out/prof_r1 python tools/one_step.py 2 > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
echo "== bench ours"; timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_ours.json
