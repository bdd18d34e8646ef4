# Round-2 ncu evidence (run on the GPU box AFTER the same commands have exited 0 without ncu).
#   1. launch list of the bench command (gpu__time_duration, no clock control)
#   2. ncu --set full of the headline kernel (tensor-map staged YUV420P -> RGB24), one launch
#   3. ncu --set full of the kernels changed this round
set -e
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-extra"
$BENCH > gpurun_out/r2_ncu_plain.json 2> gpurun_out/r2_ncu_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_yuv420p_rgb24_1080p.csv $BENCH > gpurun_out/r2_ncu_launches.log 2>&1 || tail -3 gpurun_out/r2_ncu_launches.log
OUT=gpurun_out/r2_ncu_summaries.md
echo "# ncu --set full (no clock control), one launch each, round-2 code; summarised by tools/ncu_summary.py" > $OUT
cap() {   # label, command...
  local label="$1"; shift
  "$@" > /dev/null 2>&1
  rm -f /tmp/x.ncu-rep
  ncu --set full --clock-control none --import-source on $NCUSEL -s ${SKIP:-3} -c 1 -o /tmp/x "$@" > /tmp/ncu.log 2>&1 || { echo "ncu failed for $label"; tail -3 /tmp/ncu.log; return 0; }
  echo -e "\n## $label\n\`$*\`\n\n\`\`\`" >> $OUT
  python tools/ncu_summary.py /tmp/x.ncu-rep | grep -v "^==" >> $OUT
  echo '```' >> $OUT
}
NCUSEL="-k regex:k_yuv420_rgb24_tma2d"
cap "HEADLINE: YUV420P -> RGB24 1080p, 256 frames per launch, automatic path (tensor-map staged loads, 3 stages)" $BENCH
cp /tmp/x.ncu-rep gpurun_out/r2_headline.ncu-rep 2>/dev/null || true
NCUSEL="-k regex:k_yuv2rgb"
ACGPU_TMA_AUTO=0 cap "same with ACGPU_TMA_AUTO=0: tier 2 (LDG loads)" env ACGPU_TMA_AUTO=0 $BENCH
NCUSEL=""
SKIP=3 cap "UHD round trip chain, leg 2: RGB24 -> YUV422P" python tools/sweep.py --steps 1 --size 3840x2160 --pairs rgb24:yuv422p
cap "antialias Y random bytes" python tools/tcv_probe.py --frames 32 --steps 1 --bpp 1 --only "antialias random"
cap "antialias Y gradient" python tools/tcv_probe.py --frames 32 --steps 1 --bpp 1 --only "antialias gradient"
cap "antialias RGB24 gradient" python tools/tcv_probe.py --frames 32 --steps 1 --bpp 3 --only "antialias gradient"
cap "reduce 3x3 Y" python tools/tcv_probe.py --frames 32 --steps 1 --bpp 1 --only "reduce 3x3"
cap "clip odd 3/5/1/1 Y" python tools/tcv_probe.py --frames 32 --steps 1 --bpp 1 --only "clip odd"
cap "flip_v in place RGB24" python tools/tcv_probe.py --frames 32 --steps 1 --bpp 3 --only "flip_v in place"
wc -l $OUT
