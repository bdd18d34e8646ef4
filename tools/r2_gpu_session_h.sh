python -m pytest tests/test_gpu_parity.py tests/test_gpu_helpers.py tests/test_gpu_fuzz.py tests/test_tcv_shim.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2h_tests.log
./tools/legacy_bench > gpurun_out/r2h_legacy_bench.txt 2>&1
( for v in tcv_caller tcv_caller_ref_sse2; do echo $v; ./tests/c/$v time 200; done ) > gpurun_out/r2h_tcv_time.txt 2>&1
python tools/tcv_probe.py --only reduce > gpurun_out/r2h_reduce.txt 2>&1
sh tools/r2_ncu_capture.sh > gpurun_out/r2h_ncu.log 2>&1
