#!/bin/bash
# Round 2, visit K (1 GPU): exact per-tile culling at binning -- tests, A/B bench, ncu capture of one whole step.
mkdir -p gpurun_out; rm -f gpurun_out/rowcheck.jsonl
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
for mode in 1 0; do
  GRPG_EXACT_TILE_CULL=$mode timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_cull$mode.json
  python - $mode <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_cull{sys.argv[1]}.json")); k = d["kernels"]
print("GRPG_EXACT_TILE_CULL=" + sys.argv[1], "graph step %.4f fwd %.4f | eager step %.4f fwd %.4f | e2e %.1f | binned %d" % (d["ms_per_step"], d["fwd_ms"], d.get("ms_per_step_eager", 0), d.get("fwd_ms_eager", 0), d["e2e"]["value"], d["index_check"]["binned"]))
print("   ", {n: round(v["ms_per_step"], 4) for n, v in k.items()})
print("    parity", json.dumps(d.get("parity"))[:400])
PY
done
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blend_fwd|blend_bwd|onesweep_pass|radix_tile|emit_instances|preprocess_fwd|preprocess_bwd|scan_tiles|tile_ranges" -s 24 -c 26 -f -o gpurun_out/prof_final python tools/one_step.py 3 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
python tools/summarize_ncu.py gpurun_out/prof_final.ncu-rep gpurun_out/r2_ncu_summary 2>&1 | tail -1
