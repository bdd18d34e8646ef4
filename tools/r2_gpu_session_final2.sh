# final 1-GPU session of round 2 (second half): the driver's own sequence (GPU tests, smoke, reference arm, bench) and the sweeps quoted in DESIGN.md
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2z_tests.log
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2z_smoke.log 2>&1
python bench.py --impl reference > gpurun_out/r2z_ref.json 2> gpurun_out/r2z_ref.err
python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
python tools/tcv_probe.py > gpurun_out/r2z_tcv_probe.txt 2>&1
python tools/resize_probe.py > gpurun_out/r2z_resize_probe.txt 2>&1
python tools/sweep.py --pairs cfg2 --out gpurun_out/r2z_sweep_cfg2_1080p.md > gpurun_out/r2z_sweep_cfg2.log 2>&1
python tools/sweep.py --pairs all --out gpurun_out/r2z_sweep_all225_1080p.md > gpurun_out/r2z_sweep_all.log 2>&1
( for sz in 854x480 1080x1920 766x512; do echo "## $sz"; python tools/sweep.py --size $sz --pairs yuv420p:rgb24,yuv420p:rgba32,rgb24:yuv420p,bgra32:yuv420p,yuv420p:yuv422p,yuv422p:yuv420p,yuv420p:yuy2,uyvy:yuv420p,yuv444p:yuv420p,yuv420p:yuv444p,yuv420p:yuv420p; done ) > gpurun_out/r2z_sweep_ragged_420_widths.md 2>&1
( for sz in 3840x2160 1280x720 720x576 7680x4320; do echo "## $sz"; python tools/sweep.py --size $sz --pairs yuv420p:rgb24,rgb24:yuv422p,rgb24:yuv420p,yuy2:yuv420p; done ) > gpurun_out/r2z_sweep_sizes.md 2>&1
python bench.py --workload process_frame_1080p --steps 8 > gpurun_out/r2z_process_frame.json 2>gpurun_out/r2z_process_frame.err
timeout 600 python tests/soak_fuzz.py 150 > gpurun_out/r2z_soak.log 2>&1
