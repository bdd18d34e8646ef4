python -m pytest tests/test_gpu_chain.py tests/test_gpu_knobs.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2j_tests.log
( echo "fused"; python bench.py --workload uhd_roundtrip --no-e2e --no-cpu --steps 10
  echo "ACGPU_CHAIN_FUSE=0"; ACGPU_CHAIN_FUSE=0 python bench.py --workload uhd_roundtrip --no-e2e --no-cpu --steps 10 ) > gpurun_out/r2j_fuse.txt 2>&1
