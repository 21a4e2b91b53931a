#!/bin/bash
# Round 2, visit L (1 GPU): cooperative tile cull, slot-11 backward, batched blend kernels -- tests, A/B bench, band A/B.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
for mode in 1 0; do
  GRPG_EXACT_TILE_CULL=$mode timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_cull$mode.json
  python - $mode <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_cull{sys.argv[1]}.json")); k = d["kernels"]
print("GRPG_EXACT_TILE_CULL=" + sys.argv[1], "graph step %.4f fwd %.4f | eager step %.4f fwd %.4f | e2e %.1f | binned %d" % (d["ms_per_step"], d["fwd_ms"], d.get("ms_per_step_eager", 0), d.get("fwd_ms_eager", 0), d["e2e"]["value"], d["index_check"]["binned"]))
print("   ", {n: round(v["ms_per_step"], 4) for n, v in k.items()})
print("    parity", json.dumps(d.get("parity"))[:300])
PY
done
echo "== band A/B"
timeout 300 python tools/band_ab.py 1,8 10 2>&1 | tail -2 | tee gpurun_out/band_base.jsonl
GRPG_FWD_PIPE=1 GRPG_BWD_PIPE=1 GRPG_FWD_PIPE_CFG=54 GRPG_BWD_PIPE_CFG=52 timeout 300 python tools/band_ab.py 1,8 10 2>&1 | tail -2 | tee gpurun_out/band_pipe_54_52.jsonl
GRPG_FWD_PIPE=1 GRPG_BWD_PIPE=1 GRPG_FWD_PIPE_CFG=82 GRPG_BWD_PIPE_CFG=62 timeout 300 python tools/band_ab.py 1,8 10 2>&1 | tail -2 | tee gpurun_out/band_pipe_82_62.jsonl
GRPG_FWD_PIPE=1 GRPG_BWD_PIPE=1 GRPG_FWD_PIPE_CFG=44 GRPG_BWD_PIPE_CFG=44 timeout 300 python tools/band_ab.py 1,8 10 2>&1 | tail -2 | tee gpurun_out/band_pipe_44_44.jsonl
