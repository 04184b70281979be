#!/bin/bash
# quick GPU iteration: parity tests + splat timing at config-3 shape
set -u
OUT=gpurun_out/${1:-quick}
mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee "$OUT/pytest_gpu.log"
timeout 300 python scripts/quick_splat_time.py 64 2>&1 | tail -5 | tee "$OUT/quick.log"
