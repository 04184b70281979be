#!/bin/bash
set -u
OUT=gpurun_out/${1:-mprobe}; mkdir -p "$OUT"
run() { tag=$1; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e 2>"$OUT/$tag.err" > "$OUT/$tag.json"; python - "$OUT/$tag.json" "$tag" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); print(sys.argv[2], round(d["value"]), round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d["roofline"]["kernels_ms"].items()}, d["multi_gpu_check"]["exchange"] if d.get("multi_gpu_check") else None)
except Exception as e: print(sys.argv[2], "ERR", e)
PY
}
run noself_nccl FFB_BENCH_SKIP_SELFCHECK=1 FFB_SYMM_ALLREDUCE=0
run noself_peer FFB_BENCH_SKIP_SELFCHECK=1
run self_nccl FFB_SYMM_ALLREDUCE=0
echo "== pytest multi gpu"
timeout 240 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -15
