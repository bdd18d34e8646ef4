# clip (two-phase window kernel) and gamma (private tables) A/B against the previous build, then their parity tests
O=gpurun_out/r2p_tcv_ab.txt; : > $O
for lib in libacgpu_base.so libacgpu.so; do
  echo "## $lib" >> $O
  ACGPU_LIB=$PWD/transcode-tcforge_b200/$lib python tools/tcv_probe.py --only clip >> $O 2>&1
  ACGPU_LIB=$PWD/transcode-tcforge_b200/$lib python tools/tcv_probe.py --only "reduce 1x2" >> $O 2>&1
  ACGPU_LIB=$PWD/transcode-tcforge_b200/$lib python tools/tcv_probe.py --only gamma >> $O 2>&1
done
for wv in 1 4 8; do echo "## libacgpu.so ACGPU_WAVES=$wv (gamma)" >> $O; ACGPU_WAVES=$wv python tools/tcv_probe.py --only gamma >> $O 2>&1; done
python -m pytest tests/test_gpu_tcvops.py tests/test_gpu_chain.py tests/test_tcv_shim.py tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2p_tests.log
