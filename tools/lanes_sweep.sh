#!/bin/bash
# development: timed-region throughput against the number of stream lanes
mkdir -p gpurun_out
for l in 1 2 3 4 6; do
  SHARDMERGE_LANES=$l timeout 600 python bench.py --layers 4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lanes $l value %.2f G ms %.2f'%(d['value']/1e9,d['ms_per_step']))"
done
