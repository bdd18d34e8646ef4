python -m pytest tests/test_gpu_tcvops.py tests/test_gpu_chain.py tests/test_tcv_shim.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2c_tests.log
python tools/tcv_probe.py > gpurun_out/r2c_tcv_probe.txt 2>&1
for u in 1 2 4; do echo "ACGPU_FLIPV_UNROLL=$u"; ACGPU_FLIPV_UNROLL=$u python tools/tcv_probe.py --only flip_v; done > gpurun_out/r2c_flipv.txt 2>&1
