#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
echo "== bench ours"; timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_ours.json
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
