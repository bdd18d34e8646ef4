# flat tensor-map staged loads (k_yuv420_rgb24_tmaflat): parity first, then sizes x {off, widths that do not fill warps, every width}, block 128 / 256
python -m pytest tests/test_gpu_parity.py tests/test_gpu_knobs.py tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2cc_tests.log
O=gpurun_out/r2cc_tmaflat.txt; : > $O
for sz in 720x576 1280x720 640x480 800x600 1600x900 1920x1080 3840x2160 2560x1440; do
  for mode in 0 2; do
    echo "## $sz ACGPU_TMA_FLAT=$mode" >> $O
    ACGPU_TMA_FLAT=$mode python tools/sweep.py --size $sz --pairs yuv420p:rgb24 >> $O 2>&1
  done
  echo "## $sz ACGPU_TMA_FLAT=2 block 256" >> $O
  ACGPU_TMA_FLAT=2 ACGPU_TMA_FLAT_BLOCK=256 python tools/sweep.py --size $sz --pairs yuv420p:rgb24 >> $O 2>&1
  echo "## $sz ACGPU_TMA_FLAT=2 block 64" >> $O
  ACGPU_TMA_FLAT=2 ACGPU_TMA_FLAT_BLOCK=64 python tools/sweep.py --size $sz --pairs yuv420p:rgb24 >> $O 2>&1
done
