#!/bin/bash
# A-B of library builds on the Llama-8B-shaped bench (8 layers resident): tools/ab_lib.sh "" build_tmp/lib_x.so ...   ("" = in-tree)
for lib in "$@"; do
  echo "== lib=${lib:-in-tree}"
  if [ -n "$lib" ]; then export SHARDMERGE_B200_LIB=$PWD/$lib; else unset SHARDMERGE_B200_LIB; fi
  python bench.py --layers 8 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1])
pc=l['roofline']['per_class']
print('value %.2f Gparam/s  serial %.2f  col frac %.3f' % (l['value']/1e9, l['roofline']['serial_pass']['params_per_s_per_gpu']/1e9, l['roofline']['frac']))
for k in ('row_fwd','col_fwd','stats_cutoff','blend_cull','col_inv','row_inv'):
    print('  %-13s %.2f ms/step  %.0f GB/s' % (k, pc[k]['ms']/l['steps'], pc[k]['gbs']))
"
done
unset SHARDMERGE_B200_LIB
