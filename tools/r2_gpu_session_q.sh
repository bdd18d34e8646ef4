# window kernel with loads issued before shifts (main: 64 registers; w5: min 5 blocks, 48 registers), gamma unroll 2 (main) vs 4 (w5)
O=gpurun_out/r2q_tcv_ab.txt; : > $O
for lib in libacgpu_base.so libacgpu.so libacgpu_w5.so; do
  echo "## $lib" >> $O
  ACGPU_LIB=$PWD/transcode-tcforge_b200/$lib python tools/tcv_probe.py --only clip >> $O 2>&1
  ACGPU_LIB=$PWD/transcode-tcforge_b200/$lib python tools/tcv_probe.py --only "reduce 1x2" >> $O 2>&1
  ACGPU_LIB=$PWD/transcode-tcforge_b200/$lib python tools/tcv_probe.py --only gamma >> $O 2>&1
done
python -m pytest tests/test_gpu_tcvops.py tests/test_gpu_chain.py tests/test_tcv_shim.py tests/test_gpu_fuzz.py tests/test_frame_plumbing.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2q_tests.log
