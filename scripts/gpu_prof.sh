#!/bin/bash
# ncu full capture of the splat kernels on the config-3 shape (B=64): gpurun -- bash scripts/gpu_prof.sh <tag> [kernel regex]
set -u
OUT=gpurun_out/${1:-prof}
KR=${2:-splat_|prepare}
mkdir -p "$OUT"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$KR" -s 8 -c 4 -f -o "$OUT/prof" \
    python scripts/quick_splat_time.py 64 > "$OUT/ncu.log" 2>&1
tail -3 "$OUT/ncu.log"
