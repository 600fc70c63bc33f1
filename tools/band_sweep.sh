#!/bin/bash
# A-B: column band size (SM_COL_BAND_MB) on the Llama-8B-shaped bench, 4 layers resident
for mb in 0 16 32 64; do
  echo "== SM_COL_BAND_MB=$mb"
  SM_COL_BAND_MB=$mb python bench.py --layers 4 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --profile-json gpurun_out/r02_band_$mb.json | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
pc=l['roofline']['per_class']
print('value %.2f Gparam/s  serial %.2f' % (l['value']/1e9, l['roofline']['serial_pass']['params_per_s_per_gpu']/1e9))
for k in ('row_fwd','col_fwd','stats_cutoff','blend_cull','col_inv','row_inv'):
    print('  %-13s %.2f ms/step  %.0f GB/s' % (k, pc[k]['ms']/l['steps'], pc[k]['gbs']))
"
done
