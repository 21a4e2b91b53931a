#!/bin/bash
# Round 2, visit A: parity tests of the packed blend kernels + A/B of the blend variants + H2D probe.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
echo "== h2d probe"; python - <<'PY' 2>&1 | tail -5
import torch, time
x = torch.empty(3, 1280, 1920).pin_memory(); d = torch.empty_like(x, device="cuda")
u = torch.empty(1280, 1920, 3, dtype=torch.uint8).pin_memory(); du = torch.empty_like(u, device="cuda")
for name, s, t in (("f32 29.5MB", x, d), ("u8 7.4MB", u, du)):
    for _ in range(3): t.copy_(s, non_blocking=True)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): t.copy_(s, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(name, "H2D ms", round(ms, 4), "GB/s", round(s.numel() * s.element_size() / ms / 1e6, 1))
PY
run() { # label, env...
  lab=$1; shift
  env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_$lab.json
  python - "$lab" <<'PY'
import json, sys
lab = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/bench_{lab}.json"))
    k = d["kernels"]
    print(lab, "step_ms %.4f fwd_ms %.4f e2e %.1f it/s | blend_fwd %.4f blend_bwd %.4f" % (d["ms_per_step"], d["fwd_ms"], d["e2e"]["value"], k["blend_fwd"]["ms_per_step"], k["blend_bwd"]["ms_per_step"]))
except Exception as e:
    print(lab, "FAILED", e, open(f"gpurun_out/bench_{lab}.json").read()[-500:])
PY
}
echo "== bench A/B"
run packed GRPG_FWD_PPL=3 GRPG_BWD_PPL=3
run scalar2 GRPG_FWD_PPL=2 GRPG_BWD_PPL=2
run packed_f6b6 GRPG_FWD_MINB=6 GRPG_BWD_MINB=6
run packed_f10b8 GRPG_FWD_MINB=10 GRPG_BWD_MINB=8
echo "== kernels (packed)"; python -c "
import json; d=json.load(open('gpurun_out/bench_packed.json')); print(json.dumps(d['kernels'])); print(d.get('roofline'))"
