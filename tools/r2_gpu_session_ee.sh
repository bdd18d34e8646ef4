# fused conversion pair behind the tensor-map staged loads: plain loads / 2 stages / 3 stages, UHD and 1080p; then the chain tests
O=gpurun_out/r2ee_fused_tma.txt; : > $O
for rep in 1 2; do for m in 0 2 3; do
  echo "## ACGPU_FUSED_TMA=$m" >> $O
  ACGPU_FUSED_TMA=$m python bench.py --workload uhd_roundtrip --no-cpu --no-e2e --no-extra --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  uhd_roundtrip', d['value'], 'frames/s  frac(unique)', d['roofline']['frac'])" >> $O
done; done
for m in 0 2 3; do
  ACGPU_FUSED_TMA=$m python -m pytest tests/test_gpu_chain.py -m gpu -q -x 2>&1 | tail -2 >> gpurun_out/r2ee_tests.log
done
