# closing 1-GPU session of round 2: the driver's own sequence on the final code, the size sweep with the flat tensor-map form, and its ncu
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2f3_tests.log
python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2f3_smoke.log 2>&1
python bench.py --impl reference > gpurun_out/r2f3_ref.json 2> gpurun_out/r2f3_ref.err
python bench.py > gpurun_out/r2f3_bench.json 2> gpurun_out/r2f3_bench.err
( for sz in 3840x2160 1920x1080 1280x720 720x576 640x480 7680x4320; do echo "## $sz"; python tools/sweep.py --size $sz --pairs yuv420p:rgb24,yuv420p:bgr24,rgb24:yuv422p,rgb24:yuv420p,yuy2:yuv420p; done ) > gpurun_out/r2f3_sweep_sizes.md 2>&1
python tools/sweep.py --pairs cfg2 --out gpurun_out/r2f3_sweep_cfg2_1080p.md > /dev/null 2>&1
OUT=gpurun_out/r2f3_ncu_summaries.md
echo "# ncu --set full (no clock control), one launch each; summarised by tools/ncu_summary.py" > $OUT
cap() {   # label, command...
  local label="$1"; shift
  "$@" > /dev/null 2>&1
  rm -f /tmp/x.ncu-rep
  ncu --set full --clock-control none --import-source on $NCUSEL -s ${SKIP:-3} -c 1 -o /tmp/x "$@" > /tmp/ncu.log 2>&1 || { echo "ncu failed for $label"; tail -3 /tmp/ncu.log; return 0; }
  echo -e "\n## $label\n\`$*\`\n\n\`\`\`" >> $OUT
  python tools/ncu_summary.py /tmp/x.ncu-rep | grep -v "^==" >> $OUT
  echo '```' >> $OUT
}
NCUSEL="-k regex:k_yuv420_rgb24_tmaflat"
SKIP=3 cap "YUV420P -> RGB24 PAL 720x576, flat tensor-map form (2 stages, blocks of 128)" python tools/sweep.py --steps 1 --size 720x576 --pairs yuv420p:rgb24
SKIP=3 cap "YUV420P -> RGB24 1280x720, flat tensor-map form" python tools/sweep.py --steps 1 --size 1280x720 --pairs yuv420p:rgb24
cuobjdump -sass transcode-tcforge_b200/libacgpu.so | grep -oE "UTMALDG[.A-Z0-9]*|UTMASTG[.A-Z0-9]*|UBLKCP[.A-Z0-9]*|SYNCS[.A-Z0-9]*" | sort | uniq -c > gpurun_out/r2f3_sass_tma_opcodes.txt
