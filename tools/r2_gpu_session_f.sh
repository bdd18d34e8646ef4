python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2f_tests.log
( echo "## default (auto)"; python tools/sweep.py --pairs yuv420p:rgb24,yuv420p:bgr24,yuv420p:rgba32,yuv420p:argb32,yuv444p:rgb24
  echo "## auto off"; ACGPU_TMA_AUTO=0 python tools/sweep.py --pairs yuv420p:rgb24,yuv420p:rgba32
  echo "## TMA=8"; ACGPU_TMA=8 python tools/sweep.py --tier 3 --pairs yuv420p:rgb24
  echo "## 720p default"; python tools/sweep.py --size 1280x720 --pairs yuv420p:rgb24,yuv420p:rgba32
  echo "## UHD default"; python tools/sweep.py --size 3840x2160 --pairs yuv420p:rgb24
  echo "## 444 smooth"; python tools/sweep.py --pairs yuv444p:rgb24 --content smooth
) > gpurun_out/r2f_sweeps.txt 2>&1
python bench.py --no-extra > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
