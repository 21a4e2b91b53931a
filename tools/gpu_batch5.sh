#!/bin/bash
mkdir -p gpurun_out
echo "== dist check (2 ranks)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 2>&1 | tail -6
echo "== bench 2 gpus"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_ours_2gpu.json
echo "== bench ref 2 gpus"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-400
