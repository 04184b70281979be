#!/bin/bash
# compute-sanitizer over the GPU parity tests: memcheck on every test file, racecheck (shared-memory hazards) and
# synccheck on the parity files except the full-size configs.  Usage: gpurun --timeout 600 -- bash scripts/gpu_sanitize.sh [tag]
set -u
OUT=gpurun_out/${1:-sanitize}
mkdir -p "$OUT"
SMALL="tests/test_curve_gpu.py tests/test_scene_gpu.py tests/test_post_gpu.py tests/test_lines_gpu.py tests/test_perlin_gpu.py tests/test_splat_gpu.py"
run() {   # tool, log name, test files
    timeout ${SAN_TIMEOUT:-200} compute-sanitizer --tool "$1" --error-exitcode 7 --print-limit 5 \
        python -m pytest $3 -m gpu -q -p no:cacheprovider > "$OUT/$2.log" 2>&1
    echo "$1 exit code: $?" | tee -a "$OUT/$2.log"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" "$OUT/$2.log" | tail -4
}
run memcheck memcheck "$SMALL tests/test_configs_gpu.py"
run racecheck racecheck "$SMALL"
run synccheck synccheck "$SMALL"
