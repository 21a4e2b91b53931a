#!/bin/bash
# one GPU-box visit: sanitizer check, golden vectors from the reference, GPU tests, side-by-side check, bench
mkdir -p gpurun_out
echo "== sanitizer S=3"; compute-sanitizer --tool memcheck --print-limit 3 python tools/small_fwd_bwd.py 3 2>&1 | grep -vE "Host Frame|^=========\s+in " | tail -15
echo "== golden"; python tests/golden/make_golden.py 2>&1 | tail -12
mkdir -p tests/golden; cp gpurun_out/golden/*.npz tests/golden/ 2>/dev/null
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
echo "== gpu_check"; timeout 600 python tools/gpu_check.py full 2>&1 | tail -12
echo "== bench ours"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_ours.json
echo "== bench ref"; timeout 600 python bench.py --impl reference --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_ref.json
