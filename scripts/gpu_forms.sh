#!/bin/bash
# backward launch forms inside the bench step: one-shot, persistent (work counter), chunks of 4 / 8 / 16 consecutive items per CTA
for rep in 1 2; do
for form in "X=1" "FFB_SPLAT_BWD_PERSIST=1" "FFB_SPLAT_BWD_PERSIST=1 FFB_SPLAT_BWD_CHUNK=4" "FFB_SPLAT_BWD_PERSIST=1 FFB_SPLAT_BWD_CHUNK=8" "FFB_SPLAT_BWD_PERSIST=1 FFB_SPLAT_BWD_CHUNK=16"; do
  echo "== $form"
  env $form timeout 300 python bench.py --no-cpu-baseline --no-side --no-e2e --steps ${STEPS:-30} 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), d['roofline']['frac'], {k: round(v,3) for k,v in d['roofline']['kernels_ms'].items()}, d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done; done
