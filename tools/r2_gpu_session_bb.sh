# fused conversion pair at 2 / 3 (shipped) / 4 blocks per SM; block width of the row-pair kernels at UHD ($ACGPU_BLOCK420)
O=gpurun_out/r2bb_fused_occ.txt; : > $O
for lib in libacgpu_fb2.so libacgpu.so libacgpu_fb4.so libacgpu_fb2.so libacgpu.so libacgpu_fb4.so; do
  echo "## $lib" >> $O
  ACGPU_LIB=$PWD/transcode-tcforge_b200/$lib python bench.py --workload uhd_roundtrip --no-cpu --no-e2e --no-extra --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  uhd_roundtrip', d['value'], 'frames/s  frac(unique)', d['roofline']['frac'])" >> $O
done
for b in 256 128 192 256 128 192; do
  echo "## ACGPU_BLOCK420=$b" >> $O
  ACGPU_BLOCK420=$b python tools/sweep.py --size 3840x2160 --pairs yuv420p:rgb24,yuv420p:yuv422p,yuy2:yuv420p >> $O 2>&1
  ACGPU_BLOCK420=$b python bench.py --workload uhd_roundtrip --no-cpu --no-e2e --no-extra --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  uhd_roundtrip', d['value'], 'frames/s  frac(unique)', d['roofline']['frac'])" >> $O
done
