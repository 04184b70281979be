#!/bin/bash
# multi-GPU bench under torchrun: gpurun --gpus N -- bash scripts/gpu_scale.sh N [tag]
N=${1:-2}; OUT=gpurun_out/${2:-scale}; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv | tee $OUT/gpus_$N.csv
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 2>$OUT/bench_$N.err | tee $OUT/bench_$N.json
tail -3 $OUT/bench_$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>>$OUT/bench_$N.err | tail -1 | cut -c1-300
