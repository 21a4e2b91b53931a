#!/bin/bash
# Round 2, visit B: full GPU test suite (incl. the config #4 training test), both bench arms with the new bench keys,
# ncu --set full of the packed blend kernels.
mkdir -p gpurun_out; rm -f gpurun_out/rowcheck.jsonl
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== bench ours"; timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_ours.json; python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_ours.json"))
for k in ("value", "ms_per_step", "ms_per_step_stats", "fwd_fps", "fwd_ms_stats", "e2e", "parity", "index_check", "roofline", "cpu_baseline", "clocks"):
    print(k, json.dumps(d.get(k))[:700])
PY
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_ref.json; python -c "
import json; d=json.load(open('gpurun_out/bench_ref.json')); print({k: d.get(k) for k in ('value','ms_per_step','fwd_fps','e2e')})"
echo "== ncu full (packed blends)"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"blend_fwd_packed|blend_bwd_packed" -s 2 -c 2 -f -o gpurun_out/prof_packed python tools/one_step.py 2 > gpurun_out/ncu_packed.log 2>&1; tail -2 gpurun_out/ncu_packed.log
python tools/summarize_ncu.py gpurun_out/prof_packed.ncu-rep gpurun_out/r2_ncu_packed 2>&1 | tail -1; cat gpurun_out/r2_ncu_packed.json | head -80
