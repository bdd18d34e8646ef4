python -m pytest tests/test_gpu_tcvops.py tests/test_gpu_chain.py tests/test_tcv_shim.py tests/test_gpu_fuzz.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2m_tests.log
python tools/tcv_probe.py --only antialias > gpurun_out/r2m_aa.txt 2>&1
python tools/tcv_probe.py --only antialias --frames 8 >> gpurun_out/r2m_aa.txt 2>&1
