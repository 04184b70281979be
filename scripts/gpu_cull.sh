for L in 20.7233 18.42 16.118 13.8155; do
  echo "== ln(1/tau) = $L"
  FFB_BWD_CULL_LN=$L timeout 300 python scripts/grad_precision.py 2>&1 | grep -E "st rebuild  |st rebuild *:" | head -2
  FFB_BWD_CULL_LN=$L timeout 300 python scripts/quick_splat_time.py 64 2>&1 | grep -E "^B=64|fused L1 backward \(st\)"
done
