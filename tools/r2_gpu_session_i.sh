python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_knobs.py tests/test_gpu_chain.py tests/test_gpu_tcvops.py tests/test_frame_plumbing.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2i_tests.log
P=yuv422p:rgb24,yuv444p:rgb24,yuv411p:rgb24,yuy2:rgb24,uyvy:bgr24,yvyu:rgb24
( echo "## default (tensor-map staged loads for every YUV source -> RGB24/BGR24)"; python tools/sweep.py --pairs $P
  echo "## ACGPU_TMA_AUTO=1 (4:2:0 only: the others on tier 2)"; ACGPU_TMA_AUTO=1 python tools/sweep.py --pairs $P
  echo "## 720p default"; python tools/sweep.py --size 1280x720 --pairs $P
  echo "## 720p ACGPU_TMA_AUTO=1"; ACGPU_TMA_AUTO=1 python tools/sweep.py --size 1280x720 --pairs $P
  echo "## smooth content default"; python tools/sweep.py --pairs yuv444p:rgb24,yuy2:rgb24 --content smooth
  echo "## smooth content ACGPU_TMA_AUTO=1"; ACGPU_TMA_AUTO=1 python tools/sweep.py --pairs yuv444p:rgb24,yuy2:rgb24 --content smooth
) > gpurun_out/r2i_sweeps.txt 2>&1
( for sz in 854x480 1080x1920 766x512; do echo "## $sz"; python tools/sweep.py --size $sz --pairs yuv420p:rgb24,yuv420p:rgba32,rgb24:yuv420p,yuv420p:yuv422p,yuv422p:yuv420p,yuv420p:yuy2,uyvy:yuv420p,yuv444p:yuv420p,yuv420p:yuv444p; done ) > gpurun_out/r2i_ragged.txt 2>&1
python tools/tcv_probe.py --only clip > gpurun_out/r2i_clip.txt 2>&1
./tools/legacy_bench > gpurun_out/r2i_legacy_bench.txt 2>&1
