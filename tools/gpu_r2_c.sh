#!/bin/bash
# Round 2, visit C: GPU tests (static-capacity mode, compacting depth sort, training trajectory) + bench with the graph mode.
mkdir -p gpurun_out; rm -f gpurun_out/rowcheck.jsonl
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== bench ours"; timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_ours.json; python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_ours.json"))
for k in ("value", "ms_per_step", "value_eager", "ms_per_step_eager", "ms_per_step_stats", "fwd_fps", "e2e", "cuda_graph", "execution", "kernels", "roofline", "cpu_baseline"):
    print(k, json.dumps(d.get(k))[:900])
PY
