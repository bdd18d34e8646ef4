# horizontal resize: rewritten window kernel (libacgpu.so) against the previous commit (libacgpu_prev.so); PAL flat walk at 5 blocks per SM (flat5); tests
O=gpurun_out/r2x_resize_ab.txt; : > $O
for lib in libacgpu_prev.so libacgpu.so libacgpu_prev.so libacgpu.so; do
  echo "## $lib" >> $O
  ACGPU_LIB=$PWD/transcode-tcforge_b200/$lib python tools/resize_probe.py >> $O 2>&1
done
for lib in libacgpu.so libacgpu_flat5.so libacgpu.so libacgpu_flat5.so; do
  echo "## $lib" >> $O
  ACGPU_LIB=$PWD/transcode-tcforge_b200/$lib python tools/sweep.py --size 720x576 --pairs yuv420p:rgb24,yuv420p:bgr24 >> $O 2>&1
  ACGPU_LIB=$PWD/transcode-tcforge_b200/$lib python tools/sweep.py --size 640x480 --pairs yuv420p:rgb24 >> $O 2>&1
done
python -m pytest tests/test_gpu_rowops.py tests/test_gpu_tcvops.py tests/test_gpu_chain.py tests/test_tcv_shim.py tests/test_gpu_fuzz.py tests/test_gpu_next_rows.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2x_tests.log
