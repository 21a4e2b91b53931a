#!/bin/bash
mkdir -p gpurun_out
echo "== pytest compose"; timeout 900 python -m pytest tests/test_compose_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -25
echo "== bench compose"; timeout 600 python tools/bench_compose.py 2>&1 | tail -3 | tee gpurun_out/compose_bench.json
