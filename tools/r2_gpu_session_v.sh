# vectorised reduce for ratios 5..8, then the plane-operation tests
python tools/tcv_probe.py --only reduce > gpurun_out/r2v_reduce.txt 2>&1
python -m pytest tests/test_gpu_tcvops.py tests/test_gpu_chain.py tests/test_tcv_shim.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2v_tests.log
