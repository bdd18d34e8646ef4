# 8-pixel row-pair kernel for width % 16 == 8 (k_yuv420_rgb24_w8) against the flat one-row units it replaces ($ACGPU_W8=0), then its tests
O=gpurun_out/r2u_w8.txt; : > $O
for sz in 1080x1920 1080x1080 360x640 600x800; do
  for w8 in 0 1; do
    echo "## $sz ACGPU_W8=$w8" >> $O
    ACGPU_W8=$w8 python tools/sweep.py --size $sz --pairs yuv420p:rgb24,yuv420p:bgr24 >> $O 2>&1
  done
done
python -m pytest tests/test_gpu_parity.py tests/test_gpu_knobs.py tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2u_tests.log
