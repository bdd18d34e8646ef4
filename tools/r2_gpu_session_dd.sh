# flat tensor-map form: 2 / 3 (shipped) / 4 stages per warp, blocks of 96 / 128 / 160 threads; then the whole suite on the final selection rule
O=gpurun_out/r2dd_tmaflat_stages.txt; : > $O
for lib in libacgpu_s2.so libacgpu.so libacgpu_s4.so; do
  for blk in 96 128 160; do
    echo "## $lib block $blk" >> $O
    for sz in 720x576 1280x720; do
      ACGPU_LIB=$PWD/transcode-tcforge_b200/$lib ACGPU_TMA_FLAT_BLOCK=$blk python tools/sweep.py --size $sz --pairs yuv420p:rgb24 >> $O 2>&1
    done
  done
done
python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2dd_tests.log
