set -e
OUT=gpurun_out/r1c_ncu_secondary.md
echo "# ncu --set full (no clock control), one launch each, final round-1 code; 1920x1080; summarised by tools/ncu_summary.py" > $OUT
cap() {   # label, command...
  local label="$1"; shift
  "$@" > /dev/null 2>&1
  rm -f /tmp/x.ncu-rep
  ncu --set full --clock-control none $NCUSEL -s 3 -c 1 -o /tmp/x "$@" > /tmp/ncu.log 2>&1 || { echo "ncu failed for $label"; tail -3 /tmp/ncu.log; return 0; }
  echo -e "\n## $label\n\`$*\`\n\n\`\`\`" >> $OUT
  python tools/ncu_summary.py /tmp/x.ncu-rep | grep -v "^==" >> $OUT
  echo '```' >> $OUT
}
NCUSEL=""
for p in rgb24:yuv420p yuv444p:rgb24 yuy2:yuv420p yuv420p:yuv422p; do cap "convert $p" python tools/sweep.py --steps 1 --pairs $p; done
NCUSEL="-k regex:k_resize"
cap "horizontal resize 1920->1280 Y (window kernel)" python tools/resize_probe.py
NCUSEL=""
cap "clip crop 16/16/8/8 Y" python tools/tcv_probe.py --frames 32 --steps 1 --bpp 1 --only "clip -16"
cap "clip odd 3/5/1/1 Y" python tools/tcv_probe.py --frames 32 --steps 1 --bpp 1 --only "clip odd"
cap "reduce 2x2 Y" python tools/tcv_probe.py --frames 32 --steps 1 --bpp 1 --only "reduce 2x2"
cap "flip_v RGB24" python tools/tcv_probe.py --frames 32 --steps 1 --bpp 3 --only "flip_v"
cap "flip_h RGB24" python tools/tcv_probe.py --frames 32 --steps 1 --bpp 3 --only "flip_h"
cap "gamma Y" python tools/tcv_probe.py --frames 32 --steps 1 --bpp 1 --only "gamma"
cap "antialias Y random bytes" python tools/tcv_probe.py --frames 32 --steps 1 --bpp 1 --only "antialias random"
cap "antialias Y gradient" python tools/tcv_probe.py --frames 32 --steps 1 --bpp 1 --only "antialias gradient"
cap "antialias RGB24 gradient" python tools/tcv_probe.py --frames 32 --steps 1 --bpp 3 --only "antialias gradient"
wc -l $OUT
