python -m pytest tests/test_gpu_knobs.py tests/test_gpu_tcvops.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2d_tests.log
( echo "## tier 2 (default)"; python tools/sweep.py --pairs yuv420p:rgb24,yuv444p:rgb24,yuv444p:bgr24,yuv444p:rgba32,yuv411p:rgb24
  echo "## 444 smooth content"; python tools/sweep.py --pairs yuv444p:rgb24 --content smooth
  for m in 0 1 2 3 4; do echo "## tier 3 ACGPU_TMA=$m"; ACGPU_TMA=$m python tools/sweep.py --tier 3 --pairs yuv420p:rgb24; done
  echo "## 720p"; python tools/sweep.py --size 1280x720 --pairs yuv420p:rgb24
  for m in 3 4; do echo "## 720p tier 3 ACGPU_TMA=$m"; ACGPU_TMA=$m python tools/sweep.py --size 1280x720 --tier 3 --pairs yuv420p:rgb24; done
) > gpurun_out/r2d_sweeps.txt 2>&1
python tools/tcv_probe.py --only antialias > gpurun_out/r2d_aa.txt 2>&1
