#!/bin/bash
# N-GPU box: multi-GPU parity tests + bench at N ranks.   gpurun --gpus N -- bash scripts/gpu_multi.sh <tag> <N>
set -u
TAG=${1:-multi}; N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=index,name --format=csv > "$OUT/gpu.csv" 2>&1
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -15 | tee "$OUT/pytest_multi_gpu.log"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 \
    2>"$OUT/bench_n$N.err" | tee "$OUT/bench_n$N.json" | cut -c1-400
tail -3 "$OUT/bench_n$N.err"
FFB_SYMM_ALLREDUCE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e \
    2>"$OUT/bench_n${N}_nccl.err" | tee "$OUT/bench_n${N}_nccl.json" | cut -c1-200
