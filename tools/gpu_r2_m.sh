#!/bin/bash
# Round 2, visit M (1 GPU): split blend layout (single-warp CTAs), tile mask compiled out by default, pipelined emission.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_m.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_m.json")); k = d["kernels"]
print("default: graph step %.4f fwd %.4f | eager step %.4f fwd %.4f | e2e %.1f | binned %d" % (d["ms_per_step"], d["fwd_ms"], d.get("ms_per_step_eager", 0), d.get("fwd_ms_eager", 0), d["e2e"]["value"], d["index_check"]["binned"]))
print("   ", {n: round(v["ms_per_step"], 4) for n, v in k.items()})
print("    parity", json.dumps(d.get("parity"))[:300])
PY
echo "== band A/B"
run() { echo "-- $1"; env $1 timeout 300 python tools/band_ab.py 1,2,4,8 10 2>&1 | tail -4 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    k = d['kernels_ms']; print(d['stride'], 'fwd %.4f bwd %.4f emit %.4f pre %.4f  sum %.4f' % (k['blend_fwd'], k['blend_bwd'], k['emit_instances'], k['preprocess_fwd'], sum(k.values())), d['checksum'])
"; }
run "GRPG_X=0" | tee gpurun_out/band_m.log
run "GRPG_BLEND_SPLIT=1" | tee -a gpurun_out/band_m.log
run "GRPG_BLEND_SPLIT=1 GRPG_FWD_PIPE=1" | tee -a gpurun_out/band_m.log
run "GRPG_FWD_PIPE=1 GRPG_FWD_PIPE_CFG=54" | tee -a gpurun_out/band_m.log
