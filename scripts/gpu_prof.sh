#!/bin/bash
# ncu full capture of the backward kernels on the config-3 shape (B=64), third iteration:
#   gpurun -- bash scripts/gpu_prof.sh <tag> [kernel regex]
set -u
OUT=gpurun_out/${1:-prof}
K=${2:-splat_bwd_st}
mkdir -p "$OUT"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$K" -s ${4:-2} -c 1 -f -o "$OUT/prof" \
    python scripts/prof_splat.py 64 ${3:-} > "$OUT/ncu.log" 2>&1
tail -2 "$OUT/ncu.log"
