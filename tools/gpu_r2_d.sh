#!/bin/bash
# Round 2, visit D (2 GPUs): tests, TMA pack A/B on one GPU, sharded operator check + bench on two.
mkdir -p gpurun_out; rm -f gpurun_out/rowcheck.jsonl
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== TMA pack A/B"; GRPG_TMA_PACK=1 timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "vs_golden or full_size" 2>&1 | tail -2
for mode in 0 1; do
  GRPG_TMA_PACK=$mode timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-graph 2>&1 | tail -1 > gpurun_out/bench_tma$mode.json
  python - $mode <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_tma{sys.argv[1]}.json")); k = d["kernels"]
print("GRPG_TMA_PACK=" + sys.argv[1], "fwd_ms %.4f step_ms %.4f" % (d["fwd_ms"], d["ms_per_step"]), {n: round(v["ms_per_step"], 4) for n, v in k.items() if "blend_fwd" in n or "pack" in n or "sort_pass" in n})
PY
done
echo "== dist check (2 GPUs)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check.log 2>&1; grep -E "output=|CUDA graph|Error|error|Traceback" gpurun_out/dist_check.log | cut -c1-400 | tail -12
echo "== dist check full (2 GPUs)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/dist_check.py full > gpurun_out/dist_check_full.log 2>&1; grep -E "output=|CUDA graph|Error|error|Traceback" gpurun_out/dist_check_full.log | cut -c1-400 | tail -12
echo "== bench 2 gpus"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench2.log 2>&1; tail -1 gpurun_out/bench2.log > gpurun_out/bench_ours_2gpu.json; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_ours_2gpu.json"))
    for k in ("value", "ms_per_step", "ms_per_step_stats", "fwd_fps", "e2e", "execution", "kernels", "collectives", "roofline", "parity", "camera_parallel"):
        print(k, json.dumps(d.get(k))[:800])
except Exception as e:
    print("bench2 FAILED", e); print(open("gpurun_out/bench2.log").read()[-3000:])
PY
