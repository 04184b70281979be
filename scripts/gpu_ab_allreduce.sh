#!/bin/bash
# A/B on one box: step time with the peer-memory fold+allreduce vs fold + NCCL: gpurun --gpus N -- bash scripts/gpu_ab_allreduce.sh N
N=${1:-4}; OUT=gpurun_out/ab_allreduce; mkdir -p $OUT
for mode in 1 0 1 0; do
  FFB_SYMM_ALLREDUCE=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$mode \
      bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-side 2>/dev/null | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('symm=$mode', d['value'], d['ms_per_step'], d['roofline']['kernels_ms']['fold'])" | tee -a $OUT/ab_$N.txt
done
