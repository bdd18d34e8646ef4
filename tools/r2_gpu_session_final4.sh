# last 1-GPU session of round 2: the driver's own sequence on the final code (tests, smoke, reference arm, bench) and the size sweep
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2f4_tests.log
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2f4_smoke.log 2>&1
python bench.py --impl reference > gpurun_out/r2f4_ref.json 2> gpurun_out/r2f4_ref.err
python bench.py > gpurun_out/r2f4_bench.json 2> gpurun_out/r2f4_bench.err
( for sz in 3840x2160 1920x1080 1280x720 720x576 640x480 7680x4320; do echo "## $sz"; python tools/sweep.py --size $sz --pairs yuv420p:rgb24,yuv420p:bgr24,rgb24:yuv422p,rgb24:yuv420p,yuy2:yuv420p; done ) > gpurun_out/r2f4_sweep_sizes.md 2>&1
timeout 300 python tests/soak_fuzz.py 1500 > gpurun_out/r2f4_soak.log 2>&1
