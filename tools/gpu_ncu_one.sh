#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${2:-2} -c ${3:-2} -f -o gpurun_out/prof_one python tools/one_step.py 2 > gpurun_out/ncu_one.log 2>&1; tail -2 gpurun_out/ncu_one.log
