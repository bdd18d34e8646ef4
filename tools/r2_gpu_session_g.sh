python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2g_tests.log
./tools/legacy_bench > gpurun_out/r2g_legacy_bench.txt 2>&1
( for v in tcv_caller tcv_caller_ref tcv_caller_ref_sse2; do echo $v; ./tests/c/$v time 200; done; echo tcv_caller_legacy; ./tests/c/tcv_caller_legacy time 20 ) > gpurun_out/r2g_tcv_time.txt 2>&1
( echo "ACGPU_COPY_THREADS=0"; ACGPU_COPY_THREADS=0 ./tools/legacy_bench 1920 1080 1.0 ) >> gpurun_out/r2g_legacy_bench.txt 2>&1
