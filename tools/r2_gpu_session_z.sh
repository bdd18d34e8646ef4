# diagnostics: the conversions INTO ragged 4:2:0 (854x480)
OUT=gpurun_out/r2zz_ncu_summaries.md
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
NCUSEL="-k regex:k_linear"
SKIP=3 cap "YUV444P -> YUV420P 854x480 (Ragged420To)" python tools/sweep.py --steps 1 --size 854x480 --pairs yuv444p:yuv420p
SKIP=3 cap "UYVY -> YUV420P 854x480 (Ragged420To)" python tools/sweep.py --steps 1 --size 854x480 --pairs uyvy:yuv420p
SKIP=3 cap "YUV422P -> YUV420P 854x480 (Ragged420To)" python tools/sweep.py --steps 1 --size 854x480 --pairs yuv422p:yuv420p
