#!/bin/bash
# Round-end measurement set: GPU tests, side-by-side parity/timing vs the reference extension, both bench arms,
# the next-row benches, ncu launch list of the bench command and one ncu --set full capture of the hot kernels.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=line 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== side by side"; timeout 600 python tools/gpu_check.py full > gpurun_out/gpu_check.log 2>&1; tail -1 gpurun_out/gpu_check.log | cut -c1-300
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_ref.json; cut -c1-300 gpurun_out/bench_ref.json
echo "== bench ours"; timeout 900 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 > gpurun_out/bench_ours.json; cut -c1-400 gpurun_out/bench_ours.json
echo "== next rows"; timeout 300 python tools/bench_compose.py 2>&1 | tail -1 > gpurun_out/compose_bench.json
timeout 300 python tools/bench_optim.py 2>&1 | tail -1 > gpurun_out/optim_bench.json
timeout 600 python tools/train_iter_bench.py 2>&1 | tail -1 > gpurun_out/train_iter.json; cut -c1-400 gpurun_out/train_iter.json
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches.log 2>&1; tail -1 gpurun_out/launches.log | cut -c1-200
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blend_fwd|blend_bwd|onesweep_pass|radix_tile|emit_instances|preprocess_fwd|preprocess_bwd|scan_tiles|tile_ranges" -s 19 -c 19 -f -o gpurun_out/prof_final python tools/one_step.py 2 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
