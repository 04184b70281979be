#!/bin/bash
# ncu full capture of the binning kernel at B=256: gpurun -- bash scripts/gpu_prof_prepare.sh <tag>
set -u
OUT=gpurun_out/${1:-prof_prep}
mkdir -p "$OUT"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:prepare" -s 2 -c 1 -f -o "$OUT/prof" \
    python scripts/prof_splat.py 256 > "$OUT/ncu.log" 2>&1
tail -2 "$OUT/ncu.log"
