# pipeline chunk schedule A/B (ramped first / last chunks vs equal chunks), end-to-end numbers of three workloads; then ncu of the changed kernels
O=gpurun_out/r2s_pipe_ramp.txt; : > $O
for wl in yuv420p_rgb24_1080p deinterlace_blend_1080p_rgb uhd_roundtrip yuy2_yuv420p_720p; do
  for ramp in 0 1 0 1; do
    echo "workload=$wl ramp=$ramp" >> $O
    ACGPU_PIPE_RAMP=$ramp python bench.py --workload $wl --no-cpu --no-extra --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('  e2e', e['value'], 'ceiling', e['ceiling']['frames_per_s'], 'frac', e['ceiling']['frac'])" >> $O
  done
done
python -m pytest tests/test_gpu_chain.py tests/test_gpu_next_rows.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/r2s_tests.log
OUT=gpurun_out/r2s_ncu_summaries.md
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
cap "clip odd 3/5/1/1 Y (k_window<shifted>: mixed chunks in a second phase, loads issued before shifts)" python tools/tcv_probe.py --frames 32 --steps 1 --bpp 1 --only "clip odd"
cap "clip -16/-16/-8/-8 Y (k_window<aligned>)" python tools/tcv_probe.py --frames 32 --steps 1 --bpp 1 --only "clip -16"
NCUSEL="-k regex:k_yuv420_rgb_yuv"
SKIP=2 cap "fused YUV420P -> RGB -> YUV422P, UHD (BASELINE config 4 as a chain), 64 frames per launch" python bench.py --workload uhd_roundtrip --steps 2 --warmup 3 --no-cpu --no-e2e --no-extra
