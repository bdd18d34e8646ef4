# fused conversion pair: components from unpacked channel words (dp2a.hi on G|R and on B's own word) against the packed-pixel form of the previous commit
O=gpurun_out/r2ff_fused_unpacked.txt; : > $O
for lib in libacgpu_prev.so libacgpu.so libacgpu_prev.so libacgpu.so; do
  echo "## $lib" >> $O
  ACGPU_LIB=$PWD/transcode-tcforge_b200/$lib python bench.py --workload uhd_roundtrip --no-cpu --no-e2e --no-extra --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  uhd_roundtrip', d['value'], 'frames/s  frac(unique)', d['roofline']['frac'])" >> $O
done
python -m pytest tests/test_gpu_chain.py -m gpu -q -x 2>&1 | tail -2 > gpurun_out/r2ff_tests.log
