# conversions INTO ragged 4:2:0: every load issued before the stores (libacgpu.so) against the previous commit (libacgpu_prev.so)
O=gpurun_out/r2aa_ragged_to.txt; : > $O
for lib in libacgpu_prev.so libacgpu.so libacgpu_prev.so libacgpu.so; do
  for sz in 854x480 1080x1920; do
    echo "## $lib $sz" >> $O
    ACGPU_LIB=$PWD/transcode-tcforge_b200/$lib python tools/sweep.py --size $sz --pairs yuv422p:yuv420p,uyvy:yuv420p,yuy2:yuv420p,yuv444p:yuv420p,yuv411p:yuv420p >> $O 2>&1
  done
done
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "ragged or fuzz or full_size" 2>&1 | tail -3 > gpurun_out/r2aa_tests.log
python -m pytest tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -3 >> gpurun_out/r2aa_tests.log
