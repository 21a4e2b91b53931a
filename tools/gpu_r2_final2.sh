#!/bin/bash
# Round 2, final single-GPU measurement set (second half of the round): GPU tests, both bench arms, ncu launch list of
# the bench command, ncu --set full captures of one whole-frame step and of the band kernels of an 8-way sharded frame.
mkdir -p gpurun_out; rm -f gpurun_out/rowcheck.jsonl
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_ref.json; cut -c1-300 gpurun_out/bench_ref.json
echo "== bench ours"; timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_ours.json; python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_ours.json"))
for k in ("value", "ms_per_step", "value_eager", "ms_per_step_stats", "fwd_fps", "fwd_fps_eager", "e2e", "parity", "roofline", "cpu_baseline", "clocks"):
    print(k, json.dumps(d.get(k))[:500])
print("kernels", {k_: round(v["ms_per_step"], 4) for k_, v in d["kernels"].items()})
PY
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/launches.log 2>&1; tail -1 gpurun_out/launches.log | cut -c1-120
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blend_fwd|blend_bwd|onesweep_pass|radix_tile|emit_instances|preprocess_fwd|preprocess_bwd|scan_tiles|tile_ranges" -s 22 -c 22 -f -o gpurun_out/prof_final python tools/one_step.py 2 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
python tools/summarize_ncu.py gpurun_out/prof_final.ncu-rep gpurun_out/r2_ncu_summary 2>&1 | tail -1
echo "== ncu band"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"blend_fwd|blend_bwd|emit_instances|preprocess_fwd" -s 4 -c 4 -f -o gpurun_out/prof_band8 python tools/one_step.py 2 2000000 8 > gpurun_out/ncu_band.log 2>&1; tail -2 gpurun_out/ncu_band.log
python tools/summarize_ncu.py gpurun_out/prof_band8.ncu-rep gpurun_out/r2_ncu_band8 2>&1 | tail -1
