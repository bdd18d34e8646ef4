# lone caller on pageable frames: piece size of the shared host copies x helper threads
O=gpurun_out/r2jj_copy_pieces.txt; : > $O
for kb in 256 128 64; do for n in 3 7 11; do
  echo "## ACGPU_COPY_PIECE_KB=$kb ACGPU_COPY_THREADS=$n" >> $O
  ACGPU_COPY_PIECE_KB=$kb ACGPU_COPY_THREADS=$n ./tools/legacy_bench 1920 1080 1.0 2>&1 | grep -E "pageable" | head -2 >> $O
done; done
