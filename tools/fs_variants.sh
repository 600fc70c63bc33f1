#!/bin/bash
# development: time the statistics passes for the in-tree library and every build_tmp/lib_*.so variant (SHARDMERGE_B200_LIB),
# plain spectra and round-2-like ones (20 % of the bins at rounding-noise level, exact zeros)
mkdir -p gpurun_out
for lib in "" $(ls build_tmp/lib_*.so 2>/dev/null); do
  echo "=== ${lib:-in-tree}"
  if [ -n "$lib" ]; then export SHARDMERGE_B200_LIB=$PWD/$lib; else unset SHARDMERGE_B200_LIB; fi
  for shp in "1024 4096" "4096 4096" "14336 4096"; do
    python tools/time_fstats.py $shp 2>&1 | tail -4 | head -3 | tr '\n' ' '; echo
    python tools/time_fstats.py $shp culled 2>&1 | tail -4 | head -3 | tr '\n' ' '; echo
  done
done
unset SHARDMERGE_B200_LIB
