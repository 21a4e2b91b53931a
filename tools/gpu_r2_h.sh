#!/bin/bash
# Round 2, visit H (8 GPUs): the sharded bench at 8 and 4 ranks (sparse exchange default, NCCL reduce-scatter for A/B at 8).
mkdir -p gpurun_out
run() { # n, label, env
  n=$1; lab=$2; shift 2
  echo "== bench $n gpus ($lab)"
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n$((RANDOM % 9)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench${n}_$lab.log 2>&1; echo "exit $?"
  grep "^{" gpurun_out/bench${n}_$lab.log | tail -1 > gpurun_out/bench_ours_${n}gpu_$lab.json
  python - $n $lab <<'PY'
import json, sys
n, lab = sys.argv[1], sys.argv[2]
try:
    d = json.load(open(f"gpurun_out/bench_ours_{n}gpu_{lab}.json"))
    for k in ("value", "ms_per_step", "ms_per_step_stats", "fwd_fps", "e2e", "execution", "collectives", "tile_load", "camera_parallel", "roofline"):
        print(k, json.dumps(d.get(k))[:500])
    print("parity", json.dumps(d["parity"]["worst_over_ranks"]))
    print("kernels", {k_: round(v["ms_per_step"], 4) for k_, v in d["kernels"].items()})
except Exception as e:
    print("FAILED", e); print(open(f"gpurun_out/bench{n}_{lab}.log").read()[-2000:])
PY
}
run 8 sparse GRPG_SPARSE_EXCHANGE=1
run 8 nccl GRPG_SPARSE_EXCHANGE=0
run 4 sparse GRPG_SPARSE_EXCHANGE=1
