python -m pytest tests/test_gpu_knobs.py tests/test_frame_plumbing.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2e_tests.log
( for sz in 1920x1080 1280x720 3840x2160 640x480; do
    echo "## $sz tier 2"; python tools/sweep.py --size $sz --pairs yuv420p:rgb24,yuv420p:bgr24
    for m in 4 5 6 7; do echo "## $sz tier 3 ACGPU_TMA=$m"; ACGPU_TMA=$m python tools/sweep.py --size $sz --tier 3 --pairs yuv420p:rgb24; done
  done
  echo "## 444"; python tools/sweep.py --pairs yuv444p:rgb24,yuv444p:bgr24,yuv444p:rgba32
  echo "## 444 smooth"; python tools/sweep.py --pairs yuv444p:rgb24 --content smooth
) > gpurun_out/r2e_sweeps.txt 2>&1
