# write-combined upload buffers: does the download rate under concurrent uploads improve?  (1 GPU)
O=gpurun_out/r2y_wc_probe.txt; : > $O
for rep in 1 2; do for mem in hostalloc wc; do for dir in h2d d2h both; do
  ./tools/host_dma_probe --gpus 1 --mem $mem --dir $dir --seconds 1.5 >> $O 2>&1
done; done; done
