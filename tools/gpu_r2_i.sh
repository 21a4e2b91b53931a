#!/bin/bash
# Round 2, visit I (8 GPUs): do the 8-warps-per-tile blend kernels (one pixel per lane) pay when a rank holds few tiles?
mkdir -p gpurun_out
run() { # n, label, env
  n=$1; lab=$2; shift 2
  echo "== bench $n gpus ($lab)"
  env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 296$n$((RANDOM % 9)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench${n}_$lab.log 2>&1; echo "exit $?"
  grep "^{" gpurun_out/bench${n}_$lab.log | tail -1 > gpurun_out/bench_ours_${n}gpu_$lab.json
  python - $n $lab <<'PY'
import json, sys
n, lab = sys.argv[1], sys.argv[2]
try:
    d = json.load(open(f"gpurun_out/bench_ours_{n}gpu_{lab}.json"))
    print("value", d["value"], "ms_per_step", d["ms_per_step"], "fwd_fps", d["fwd_fps"], "e2e", d["e2e"]["value"])
    print("parity", json.dumps(d["parity"]["worst_over_ranks"]))
    print("kernels", {k_: round(v["ms_per_step"], 4) for k_, v in d["kernels"].items()})
except Exception as e:
    print("FAILED", e); print(open(f"gpurun_out/bench{n}_{lab}.log").read()[-2000:])
PY
}
run 8 ppl1 GRPG_FWD_PPL=1 GRPG_BWD_PPL=1
run 4 ppl1 GRPG_FWD_PPL=1 GRPG_BWD_PPL=1
