#!/bin/bash
# ncu full capture of [prepare, fwd, bwd] on the config-3 shape (B=64), third iteration:
#   gpurun -- bash scripts/gpu_prof.sh <tag>
set -u
OUT=gpurun_out/${1:-prof}
mkdir -p "$OUT"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:splat_|prepare" -s 6 -c 3 -f -o "$OUT/prof" \
    python scripts/prof_splat.py 64 > "$OUT/ncu.log" 2>&1
tail -2 "$OUT/ncu.log"
