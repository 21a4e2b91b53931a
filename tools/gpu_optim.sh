#!/bin/bash
mkdir -p gpurun_out
echo "== pytest optim"; timeout 900 python -m pytest tests/test_optim_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -25
echo "== bench optim"; timeout 600 python tools/bench_optim.py 2>&1 | tail -3 | tee gpurun_out/optim_bench.json
