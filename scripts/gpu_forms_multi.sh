#!/bin/bash
# N ranks: bench step with the chunk form (default) and the one-shot form of the backward, alternating.  gpurun --gpus N -- bash scripts/gpu_forms_multi.sh N
N=${1:-4}
for rep in 1 2; do for form in "X=1" "FFB_SPLAT_BWD_PERSIST=0"; do
  echo "== $form"
  env $form FFB_BENCH_SKIP_SELFCHECK=$([ $rep = 1 ] && echo 0 || echo 1) timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$rep \
    bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-side 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), d.get('multi_gpu_check'), {k: round(v,3) for k,v in d['roofline']['kernels_ms'].items()})"
done; done
