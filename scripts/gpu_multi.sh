#!/bin/bash
# N-GPU box: multi-GPU parity tests + bench at N ranks.   gpurun --gpus N -- bash scripts/gpu_multi.sh <tag> <N>
set -u
TAG=${1:-multi}; N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=index,name --format=csv > "$OUT/gpu.csv" 2>&1
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q 2>&1 | tail -5 | tee "$OUT/pytest_multi_gpu.log"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 \
    2>"$OUT/bench_n$N.err" | tee "$OUT/bench_n$N.json" | cut -c1-200
tail -2 "$OUT/bench_n$N.err" | cut -c1-200
