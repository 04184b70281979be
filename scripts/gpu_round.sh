#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench, reference arm, ncu launch list, ncu full capture of the splat kernels.
# Usage (from the repo root):  gpurun --timeout 1500 -- bash scripts/gpu_round.sh [tag]
set -u
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/gpu.csv" 2>&1
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee "$OUT/pytest_gpu.log"
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee "$OUT/smoke.log"
echo "== bench"; timeout 900 python bench.py 2>"$OUT/bench.err" | tee "$OUT/bench.json" | cut -c1-600; tail -3 "$OUT/bench.err"
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee "$OUT/bench_ref.json" | cut -c1-300
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$OUT/launches.csv" \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-side > "$OUT/ncu_launch_bench.log" 2>&1
tail -2 "$OUT/ncu_launch_bench.log" | cut -c1-300
echo "== ncu full (splat kernels)"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:splat_(fwd_tma|bwd_st)" -s 4 -c 2 -f -o "$OUT/prof_splat" \
    python bench.py --batch 64 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-side > "$OUT/ncu_full_bench.log" 2>&1
tail -2 "$OUT/ncu_full_bench.log" | cut -c1-300
ls -la "$OUT"
