#!/bin/bash
# full ncu capture of kernels matching a regex while merging one tensor: tools/ncu_full.sh R C REGEX NAME [count]
mkdir -p gpurun_out
python tools/profile_one.py $1 $2 2 > gpurun_out/plain_one.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:$3 -c ${5:-4} -o gpurun_out/$4 -f python tools/profile_one.py $1 $2 2 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/$4.ncu-rep --page raw --csv > gpurun_out/$4_raw.csv 2>/dev/null
ls -la gpurun_out/$4*
