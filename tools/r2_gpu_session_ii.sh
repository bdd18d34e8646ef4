# lone caller on pageable frames: how many copy helper threads?
O=gpurun_out/r2ii_copy_threads.txt; : > $O
for n in 0 3 5 7 10; do
  echo "## ACGPU_COPY_THREADS=$n" >> $O
  ACGPU_COPY_THREADS=$n ./tools/legacy_bench 1920 1080 1.0 2>&1 | grep -E "pageable|threads +1 |^ *1 " | head -4 >> $O
done
