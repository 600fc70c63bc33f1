#!/bin/bash
# development: per-kernel durations and a full capture of the fused statistics kernels on one shape
mkdir -p gpurun_out
R=${1:-14336}; C=${2:-4096}
python tools/time_fstats.py $R $C > gpurun_out/fs_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/fs_launches.csv python tools/time_fstats.py $R $C > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_fs_" -s 40 -c 8 -o gpurun_out/fs_full -f python tools/time_fstats.py $R $C > gpurun_out/fs_ncu_full.log 2>&1
ncu -i gpurun_out/fs_full.ncu-rep --page raw --csv > gpurun_out/fs_full_raw.csv 2>/dev/null
cat gpurun_out/fs_plain.log
ncu -i gpurun_out/fs_full.ncu-rep --page source --csv --print-source sass > gpurun_out/fs_source.csv 2>/dev/null
