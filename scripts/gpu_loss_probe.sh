#!/bin/bash
# ncu metric pass over the fused-L1 backward (the LOSS instantiation of splat_bwd_tma) at 64 samples.
set -u
OUT=gpurun_out/${1:-lossprobe}
mkdir -p "$OUT"
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_op_read.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_executed.avg.per_cycle_active,l1tex__m_xbar2l1tex_read_bytes.sum \
    --clock-control none --kernel-name-base demangled -k 'regex:splat_bwd_tma<.*bool.1>' -c 3 --csv --log-file "$OUT/loss_metrics.csv" \
    python bench.py --batch 64 --steps 1 --warmup 1 --no-cpu-baseline > "$OUT/ncu_bench.log" 2>&1
tail -c 300 "$OUT/ncu_bench.log"
python - "$OUT/loss_metrics.csv" <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0] != "ID"]
for r in rows:
    print(r[0], r[4][:48], r[-3], r[-2], r[-1])
PY

