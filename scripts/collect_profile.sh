#!/bin/bash
# copy the evidence of one gpu_round.sh call from gpurun_out/<tag> (scratch) into profiles/<tag> (tracked):
# bench lines, test logs, the ncu launch list and a text summary of the full ncu capture.
set -u
TAG=$1
SRC=gpurun_out/$TAG; DST=profiles/$TAG
mkdir -p "$DST"
for f in bench.json bench_ref.json launches.csv pytest_gpu.log smoke.log gpu.csv; do [ -f "$SRC/$f" ] && cp "$SRC/$f" "$DST/"; done
REP=$SRC/prof_splat.ncu-rep; [ -f "$REP" ] || REP=$SRC/prof.ncu-rep
if [ -f "$REP" ]; then
  python scripts/ncu_summary.py "$REP" > "$DST/ncu_splat_summary.txt"
  python scripts/make_traffic.py "$REP" 64 "$TAG"
  python scripts/sass_summary.py > "$DST/sass_tma.txt"
  for k in fwd_tma "bwd_st$" bwd_stp bwd_tma; do
    ncu -i "$REP" --page source --csv --kernel-name regex:$k > "$SRC/src_${k//$/}.csv" 2>/dev/null
    if [ -s "$SRC/src_${k//$/}.csv" ]; then
      { echo "== $k: SASS lines bucketed by executions per super tile (warp-level), top stall lines"; python scripts/ncu_buckets.py "$SRC/src_${k//$/}.csv" ${2:-262144} 12; } >> "$DST/ncu_splat_summary.txt"
    fi
  done
fi
ls -la "$DST"
