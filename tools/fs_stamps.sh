#!/bin/bash
# development: phase timestamps of k_fs_sample / k_fs_pass for the three Llama-8B shapes
for shp in "1024 4096" "4096 4096" "14336 4096"; do
  echo "== sample $shp"; python tools/fs_stamps.py sample $shp 2>&1 | tail -7
  echo "== pass $shp"; python tools/fs_stamps.py $shp 2>&1 | tail -7
  python tools/time_fstats.py $shp 2>&1 | tail -4
done
