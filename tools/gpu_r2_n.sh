#!/bin/bash
# Round 2, visit N (1 GPU): batched backward with interleaved butterflies in the split layout -- variant tests + band A/B.
mkdir -p gpurun_out
echo "== pytest variants"; timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q --tb=short -x -k "variants or bands" 2>&1 | tail -5 | tee gpurun_out/pytest_variants.log
run() { echo "-- $1"; env $1 timeout 300 python tools/band_ab.py 1,2,4,8 10 2>&1 | tail -4 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    k = d['kernels_ms']; print(d['stride'], 'fwd %.4f bwd %.4f emit %.4f pre %.4f  sum %.4f' % (k['blend_fwd'], k['blend_bwd'], k['emit_instances'], k['preprocess_fwd'], sum(k.values())), d['checksum'])
"; }
run "GRPG_X=0" | tee gpurun_out/band_n.log
run "GRPG_BWD_PIPE=2" | tee -a gpurun_out/band_n.log
run "GRPG_BWD_PIPE=1 GRPG_BLEND_SPLIT=1" | tee -a gpurun_out/band_n.log
