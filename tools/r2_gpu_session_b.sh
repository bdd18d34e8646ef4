python -m pytest tests/test_tcv_shim.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2b_shimtest.log
for b in 50331648 104857600 209715200 419430400 1677721600 6710886400; do
  echo "ACGPU_CHAIN_L2_BYTES=$b" >> gpurun_out/r2b_chain_sweep.txt
  ACGPU_CHAIN_L2_BYTES=$b python bench.py --workload uhd_roundtrip --no-e2e --no-cpu --steps 10 >> gpurun_out/r2b_chain_sweep.txt 2>&1
  ACGPU_CHAIN_L2_BYTES=$b python bench.py --workload process_frame_1080p --no-e2e --no-cpu --steps 10 >> gpurun_out/r2b_chain_sweep.txt 2>&1
done
( for v in tcv_caller tcv_caller_ref tcv_caller_ref_sse2; do echo $v; ./tests/c/$v time 200; done; echo tcv_caller_legacy; ./tests/c/tcv_caller_legacy time 20 ) > gpurun_out/r2b_tcv_time.txt 2>&1
for mem in hostalloc register thp hugetlb; do for dir in h2d d2h both; do ./tools/host_dma_probe --gpus 1 --mem $mem --dir $dir --seconds 1; done; done > gpurun_out/r2b_dma_probe_n1.txt 2>&1
python bench.py --workload process_frame_1080p --steps 5 > gpurun_out/r2b_process_frame.json 2>gpurun_out/r2b_process_frame.err
