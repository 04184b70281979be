timeout 300 python -m pytest tests/test_splat_gpu.py tests/test_configs_gpu.py -m gpu -x -q -k "fused_l1 or launch_forms or l1_loss or sharded or shared_pattern" 2>&1 | tail -2
for rep in 1 2; do
for lib in "" "$PWD/fireflies_b200/_lib/ab/sign16.so"; do
  echo "== lib=${lib:-default(sign8)}"
  FFB_LIB=$lib timeout 300 python scripts/quick_splat_time.py 256 2>&1 | grep "fused L1 backward (st)"
  FFB_LIB=$lib timeout 200 python bench.py --no-cpu-baseline --no-side 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3))"
done; done
