#!/bin/bash
# development: parity tests touching the fused statistics + warm timings
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 400 -p no:cacheprovider -k "fused or full_size or golden or lanes or mid_size" > gpurun_out/pytest_fs.log 2>&1; tail -5 gpurun_out/pytest_fs.log
for shp in "4096 4096" "14336 4096" "1024 4096"; do echo "== $shp"; python tools/time_fstats.py $shp 2>&1 | tail -3; done
