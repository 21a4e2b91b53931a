#!/bin/bash
# Quick GPU visit: parity tests, side-by-side check with the reference extension, one bench line.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== side by side"; timeout 600 python tools/gpu_check.py full > gpurun_out/gpu_check.log 2>&1; tail -1 gpurun_out/gpu_check.log | cut -c1-1500
echo "== bench ours"; timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_ours.json; cut -c1-2500 gpurun_out/bench_ours.json
