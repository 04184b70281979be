#!/bin/bash
# fused-loss backward: counter form (default) against chunk forms, in isolation and end to end
for form in "X=1" "FFB_SPLAT_BWD_CHUNK=4" "FFB_SPLAT_BWD_CHUNK=8" "FFB_SPLAT_BWD_CHUNK=16" "FFB_SPLAT_BWD_PERSIST=0"; do
  echo "== $form"
  env $form timeout 300 python scripts/quick_splat_time.py 256 2>&1 | grep -E "fused L1 backward \(st\)"
  env $form timeout 300 python bench.py --no-cpu-baseline --no-side --steps 20 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), d['clocks']['reasons'])"
done
