timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_run.py > gpurun_out/r2_san_memcheck.log 2>&1; echo "exit $?" >> gpurun_out/r2_san_memcheck.log
ACGPU_TMA_AUTO=7 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_run.py > gpurun_out/r2_san_memcheck_tma7.log 2>&1; echo "exit $?" >> gpurun_out/r2_san_memcheck_tma7.log
