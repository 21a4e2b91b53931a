#!/bin/bash
# Round 2, visit F (2 GPUs): kNN diagnosis + band/exchange emulation tests on one GPU, then the sharded operator with the
# sparse record exchange on two (check + bench, sparse vs NCCL reduce-scatter).  Tight timeouts: a hang must not burn the budget.
mkdir -p gpurun_out
echo "== pytest subset"; timeout 600 python -m pytest tests -m gpu -q --tb=line -k "knn or band or train" 2>&1 | tail -40 | tee gpurun_out/pytest_gpu_f.log
echo "== knn diag"; timeout 300 python tools/knn_diag.py 2>&1 | tail -8 | cut -c1-600
echo "== dist check skipped"
for sp in 1 0; do
echo "== bench 2 gpus GRPG_SPARSE_EXCHANGE=$sp"; GRPG_SPARSE_EXCHANGE=$sp timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$sp bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench2_sp$sp.log 2>&1; echo "exit $?"; grep "^{" gpurun_out/bench2_sp$sp.log | tail -1 > gpurun_out/bench_ours_2gpu_sp$sp.json; python - $sp <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/bench_ours_2gpu_sp{sys.argv[1]}.json"))
    for k in ("value", "ms_per_step", "ms_per_step_stats", "fwd_fps", "e2e", "execution", "collectives", "parity", "camera_parallel"):
        print(k, json.dumps(d.get(k))[:600])
    print("kernels", {n: round(v["ms_per_step"], 4) for n, v in d["kernels"].items()})
except Exception as e:
    print("bench2 FAILED", e); print(open(f"gpurun_out/bench2_sp{sys.argv[1]}.log").read()[-2500:])
PY
done
