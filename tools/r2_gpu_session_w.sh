# diagnostics: ncu of the gamma table kernel, the horizontal resize (enlarging), the ragged 4:2:0 -> RGB24 walk and 4:4:4 -> RGB24
python tools/tcv_probe.py --only "reduce" > gpurun_out/r2w_reduce.txt 2>&1
OUT=gpurun_out/r2w_ncu_summaries.md
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
NCUSEL=""
cap "gamma 2.2 Y (k_lut)" python tools/tcv_probe.py --frames 32 --steps 1 --bpp 1 --only "gamma"
NCUSEL="-k regex:k_yuv2rgb"
SKIP=3 cap "ragged 4:2:0 -> RGB24, 1080x1920 (flat one-row units, S420R)" python tools/sweep.py --steps 1 --size 1080x1920 --pairs yuv420p:rgb24
SKIP=3 cap "YUV444P -> RGB24 1080p" python tools/sweep.py --steps 1 --size 1920x1080 --pairs yuv444p:rgb24
SKIP=3 cap "YUV420P -> RGB24 PAL 720x576 (flat walk)" python tools/sweep.py --steps 1 --size 720x576 --pairs yuv420p:rgb24
NCUSEL="-k regex:k_resize_h"
SKIP=29 cap "horizontal resize 1920 -> 2560 Y" python tools/resize_probe.py
