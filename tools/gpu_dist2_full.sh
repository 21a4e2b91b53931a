#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== dist check"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check.log 2>&1; grep -E "fused_peer|symmetric|Error|error" gpurun_out/dist_check.log | cut -c1-400 | tail -8
echo "== bench 2 gpus"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_ours_2gpu.json; cut -c1-900 gpurun_out/bench_ours_2gpu.json
