#!/bin/bash
# ncu --set full capture (with source) of the fused-L1 backward (LOSS instantiation of splat_bwd_tma) at 64 samples.
set -u
OUT=gpurun_out/${1:-lossprof}
mkdir -p "$OUT"
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:splat_bwd_tma<.*bool.1>' -c 1 -f -o "$OUT/prof_loss" \
    python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline > "$OUT/ncu_full.log" 2>&1
tail -2 "$OUT/ncu_full.log" | cut -c1-300
ls -la "$OUT"
