#!/bin/bash
# A-B of one environment switch on the Llama-8B-shaped bench (8 layers resident): tools/ab_env.sh VAR val1 val2 ...
var=$1; shift
for v in "$@"; do
  echo "== $var=$v"
  env $var=$v python bench.py --layers 8 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
pc=l['roofline']['per_class']
print('value %.2f Gparam/s  serial %.2f  col frac %.3f' % (l['value']/1e9, l['roofline']['serial_pass']['params_per_s_per_gpu']/1e9, l['roofline']['frac']))
for k in ('row_fwd','col_fwd','stats_cutoff','blend_cull','col_inv','row_inv'):
    print('  %-13s %.2f ms/step  %.0f GB/s' % (k, pc[k]['ms']/l['steps'], pc[k]['gbs']))
"
done
