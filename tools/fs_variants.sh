#!/bin/bash
# development: time the statistics passes for every build_tmp/lib_*.so variant (SHARDMERGE_B200_LIB) and the in-tree library
mkdir -p gpurun_out
for lib in "" $(ls build_tmp/lib_*.so 2>/dev/null); do
  echo "=== ${lib:-in-tree}"
  for shp in "1024 4096" "4096 4096" "14336 4096"; do
    if [ -n "$lib" ]; then export SHARDMERGE_B200_LIB=$PWD/$lib; else unset SHARDMERGE_B200_LIB; fi
    python tools/time_fstats.py $shp 2>&1 | tail -4 | tr '\n' ' '; echo
  done
done
unset SHARDMERGE_B200_LIB
