#!/bin/bash
# quick GPU iteration: splat parity tests + splat timing at config-3 shape (+ every A/B library variant under _lib/ab)
set -u
OUT=gpurun_out/${1:-quick}
mkdir -p "$OUT"
timeout 900 python -m pytest tests/test_splat_gpu.py tests/test_configs_gpu.py -m gpu -x -q 2>&1 | tail -25 | tee "$OUT/pytest_gpu.log"
timeout 300 python scripts/quick_splat_time.py 64 2>&1 | tail -16 | tee "$OUT/quick.log"
for so in fireflies_b200/_lib/ab/*.so; do
  [ -f "$so" ] || continue
  echo "== $(basename $so)"; FFB_LIB=$PWD/$so timeout 300 python scripts/quick_splat_time.py 64 2>&1 | tail -16 | tee -a "$OUT/quick.log"
done
