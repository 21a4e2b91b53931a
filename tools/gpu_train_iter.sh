#!/bin/bash
mkdir -p gpurun_out
echo "== train iteration"; timeout 900 python tools/train_iter_bench.py 2>&1 | tail -5 | tee gpurun_out/train_iter.json
