#!/bin/bash
set -u
N=${2:-4}
OUT=gpurun_out/${1:-mprobe}; mkdir -p "$OUT"
run() { tag=$1; shift; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --steps 10 --warmup 3 --no-e2e 2>"$OUT/$tag.err" > "$OUT/$tag.json"; python - "$OUT/$tag.json" "$tag" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); print(sys.argv[2], round(d["value"]), round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d["roofline"]["kernels_ms"].items()}, d["clocks"]["reasons"])
except Exception as e: print(sys.argv[2], "ERR", e)
PY
}
run persist20 X=1
run oneshot FFB_SPLAT_BWD_PERSIST=0
run old FFB_SPLAT_BWD_ST=0
run persist16 FFB_SPLAT_BWD_GRID=2368
run persist20_b X=1
run oneshot_b FFB_SPLAT_BWD_PERSIST=0
