# batch size of the do_process_frame-shaped chain, device-resident
O=gpurun_out/r2gg_chain_batch.txt; : > $O
for b in 64 128 256 384 512; do
  echo "## batch $b" >> $O
  python bench.py --workload process_frame_1080p --batch $b --no-cpu --no-e2e --no-extra --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('  process_frame_1080p', d['value'], 'frames/s  frac(unique)', r['frac'], 'sum-of-passes', r.get('sum_of_stage_passes_frac'))" >> $O
done
