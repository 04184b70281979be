#!/bin/bash
# time every A/B library variant on the config-3 splat: gpurun -- bash scripts/gpu_ab.sh
for so in fireflies_b200/_lib/ab/*.so; do
  echo "== $(basename $so)"; FFB_LIB=$PWD/$so timeout 300 python scripts/quick_splat_time.py 64 2>&1 | tail -2
done
