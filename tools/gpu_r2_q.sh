#!/bin/bash
# Round 2, last single-GPU visit: the whole GPU suite on the final code (CTA-level slot allocation in the record exchange).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
