# fused YUV420P -> RGB -> YUV kernel: dp2a on packed pixel words (f0) vs multiply-add chains on integer channels (main), and the latter at 4 blocks per SM (f4)
O=gpurun_out/r2t_fused_ab.txt; : > $O
for lib in libacgpu_f0.so libacgpu.so libacgpu_f4.so libacgpu_f0.so libacgpu.so libacgpu_f4.so; do
  echo "## $lib" >> $O
  ACGPU_LIB=$PWD/transcode-tcforge_b200/$lib python bench.py --workload uhd_roundtrip --no-cpu --no-e2e --no-extra --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  uhd_roundtrip', d['value'], 'frames/s  frac(unique)', d['roofline']['frac'])" >> $O
done
python -m pytest tests/test_gpu_chain.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/r2t_tests.log
