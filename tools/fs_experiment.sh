#!/bin/bash
# development: cost of the reductions inside the fused statistics pass (SM_FS_DBG bit0 fine, bit1 coarse, bit2 side lists)
mkdir -p gpurun_out
for shp in "4096 4096" "14336 4096"; do
  for f in 0 1 3 7; do echo "== $shp SM_FS_DBG=$f"; SM_FS_DBG=$f python tools/time_fstats.py $shp 2>&1 | tail -3; done
done
