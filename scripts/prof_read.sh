#!/bin/bash
# read a profile brought back by gpu_prof.sh: summary + per-kernel instruction buckets
D=/root/repo/gpurun_out/$1
python /root/repo/scripts/ncu_summary.py $D/prof.ncu-rep | grep -E "kernel:|duration|registers|warps_active|issue_active|inst_executed.sum|pipe_xu|dram_throughput|traffic" 
for k in fwd_tma bwd_tma prepare; do ncu -i $D/prof.ncu-rep --page source --csv --kernel-name regex:${k} > $D/src_$k.csv 2>/dev/null; done
echo FWD; python /root/repo/scripts/ncu_buckets.py $D/src_fwd_tma.csv 262144 ${2:-6}
echo BWD; python /root/repo/scripts/ncu_buckets.py $D/src_bwd_tma.csv 262144 ${2:-6}
echo PREP; python /root/repo/scripts/ncu_buckets.py $D/src_prepare.csv 2048 ${2:-6}
