# pipeline shape A/B: slots x chunk size, end-to-end numbers of three workloads
O=gpurun_out/r2l_pipe_shape.txt; : > $O
for wl in yuv420p_rgb24_1080p deinterlace_blend_1080p_rgb uhd_roundtrip; do
  for slots in 3 4 6; do for mb in 8 16 32; do
    echo "workload=$wl slots=$slots chunk_mb=$mb" >> $O
    ACGPU_PIPE_SLOTS=$slots ACGPU_PIPE_CHUNK_MB=$mb python bench.py --workload $wl --no-cpu --no-extra --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('  e2e', e['value'], 'ceiling', e['ceiling']['frames_per_s'], 'frac', e['ceiling']['frac'])" >> $O
  done; done
done
