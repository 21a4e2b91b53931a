#!/bin/bash
# Round 2, visits P/Q (N GPUs): refreshed sharded bench line with the band kernels (split layout, batched blends).
n=${1:-2}
mkdir -p gpurun_out
echo "== bench $n gpus"
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench${n}_r2b.log 2>&1; echo "exit $?"
grep "^{" gpurun_out/bench${n}_r2b.log | tail -1 > gpurun_out/bench_ours_${n}gpu_r2b.json
python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/bench_ours_{n}gpu_r2b.json"))
    print("value", d["value"], "ms_per_step", d["ms_per_step"], "fwd_fps", d["fwd_fps"], "e2e", d["e2e"]["value"])
    print("parity", json.dumps(d["parity"]["worst_over_ranks"]))
    print("kernels", {k_: round(v["ms_per_step"], 4) for k_, v in d["kernels"].items()})
    print("camera_parallel", json.dumps(d["camera_parallel"])[:300])
    print("execution", d["execution"])
except Exception as e:
    print("FAILED", e); print(open(f"gpurun_out/bench{n}_r2b.log").read()[-3000:])
PY
