#!/bin/bash
# Whole-step ncu --set full capture with the window aligned to one step (24 library kernels per step: skip the first step).
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blend_fwd|blend_bwd|onesweep_pass|radix_tile|emit_instances|preprocess_fwd|preprocess_bwd|scan_tiles|tile_ranges" -s 24 -c 24 -f -o gpurun_out/prof_final python tools/one_step.py 2 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
python tools/summarize_ncu.py gpurun_out/prof_final.ncu-rep gpurun_out/r2_ncu_summary 2>&1 | tail -1
grep -c "blend_fwd\|tile_ranges" gpurun_out/r2_ncu_summary.md
