#!/bin/bash
# e2e diagnostics: where does the host-tensor path lose against the plain pinned-copy probe?
run() { echo "== $*"; python bench.py --layers 8 --steps 3 --warmup 3 --no-cpu-baseline "$@" | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=l['e2e']
print('  e2e %.2f Gparam/s  with_files %.2f  h2d %.1f GB/s  probe %.1f' % (e['value']/1e9, e['with_files']['value']/1e9, e['h2d_gbs_achieved_rank0'], e['host_link_probe']['h2d_gbs']))"; }
run --e2e-files 0
run --e2e-files 0 --prefetch-depth 1
run --e2e-files 0 --e2e-layers 2
SHARDMERGE_LANES=1 run --e2e-files 0
run --e2e-files 1
