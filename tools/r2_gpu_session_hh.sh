# flat tensor-map form: one-segment store for warps inside one row pair (libacgpu.so) against the previous commit
O=gpurun_out/r2hh_tmaflat_store.txt; : > $O
for lib in libacgpu_prev.so libacgpu.so libacgpu_prev.so libacgpu.so; do
  echo "## $lib" >> $O
  for sz in 720x576 1280x720 640x480 1600x900; do
    ACGPU_LIB=$PWD/transcode-tcforge_b200/$lib python tools/sweep.py --size $sz --pairs yuv420p:rgb24 >> $O 2>&1
  done
done
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "flat or full_size or wider or config1" 2>&1 | tail -2 > gpurun_out/r2hh_tests.log
