#!/bin/bash
# Short validation without the ncu passes: GPU parity tests, smoke, bench, reference arm.
# Usage: gpurun --timeout 600 -- bash scripts/gpu_validate.sh [tag]
set -u
OUT=gpurun_out/${1:-validate}
mkdir -p "$OUT"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.csv" 2>&1
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee "$OUT/pytest_gpu.log"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee "$OUT/smoke.log"
timeout 600 python bench.py 2>"$OUT/bench.err" | tee "$OUT/bench.json" | cut -c1-200
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | tee "$OUT/bench_ref.json" | cut -c1-160
