#!/bin/bash
# Round 2, visit E (1 GPU): tests; the sharded bench + dist_check forced onto one rank (capture / clean-exit check);
# BASELINE config #4 at bench size; ncu launch list of the bench command.
mkdir -p gpurun_out; rm -f gpurun_out/rowcheck.jsonl
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
echo "== sharded bench forced on 1 rank"; GRPG_BENCH_FORCE_SHARDED=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_sh1.log 2>&1; echo "exit $?"; grep "^{" gpurun_out/bench_sh1.log | tail -1 > gpurun_out/bench_sh1.json; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_sh1.json"))
    for k in ("value", "ms_per_step", "fwd_fps", "e2e", "execution", "collectives", "parity"):
        print(k, json.dumps(d.get(k))[:500])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/bench_sh1.log").read()[-2500:])
PY
echo "== dist_check on 1 rank"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29522 tools/dist_check.py > gpurun_out/dist_check1.log 2>&1; echo "exit $?"; grep -E "output=|CUDA graph" gpurun_out/dist_check1.log | cut -c1-300 | tail -8
echo "== config 4 (train loop with densify/prune, both arms)"; timeout 900 python tools/train_config4.py --iters 400 2>&1 | tail -1 > gpurun_out/train_config4.json; cut -c1-3000 gpurun_out/train_config4.json
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/launches.log 2>&1; tail -1 gpurun_out/launches.log | cut -c1-200
