#!/bin/bash
# round profiles: bench (1 layer) without and with the ncu launch list, then full captures of the hot kernels
mkdir -p gpurun_out
python bench.py --layers 1 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_1layer.json 2> gpurun_out/bench_1layer.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/ncu_launches_bench_1layer.csv python bench.py --layers 1 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
python tools/profile_one.py 14336 4096 2 > gpurun_out/plain_one.log 2>&1 || exit 1
for spec in "k_col_p3:col_p3" "k_row2_inv:row2_inv"; do
  re=${spec%%:*}; nm=${spec##*:}
  ncu --set full --clock-control none --import-source on -k regex:$re -s 6 -c 4 -o gpurun_out/ncu_full_$nm -f python tools/profile_one.py 14336 4096 2 > gpurun_out/ncu_full_$nm.log 2>&1
  ncu -i gpurun_out/ncu_full_$nm.ncu-rep --page raw --csv > gpurun_out/ncu_full_${nm}_raw.csv 2>/dev/null
  rm -f gpurun_out/ncu_full_$nm.ncu-rep          # gpurun_out/ is limited to 64 MiB; the CSV export is what profiles/ keeps
done
ls -la gpurun_out/*.csv
