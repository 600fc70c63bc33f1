#!/bin/bash
# per-launch durations of one merged tensor (ncu serialises launches; cold-cache): tools/ncu_launches.sh R C
mkdir -p gpurun_out
python tools/profile_one.py $1 $2 3 > gpurun_out/plain_one.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$1x$2.csv python tools/profile_one.py $1 $2 2 > gpurun_out/ncu_one.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/launches_$1x$2.csv")))
h=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
hdr=rows[h]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); gi=hdr.index("Grid Size")
seq=[(r[ki].split("(")[0][-40:], float(r[vi].replace(",","")), r[gi]) for r in rows[h+1:] if len(r)>vi]
# last iteration only: find last k_prepare
idx=[i for i,s in enumerate(seq) if "k_row_fwd" in s[0] or "k_row2_fwd" in s[0]]
start=idx[-2] if len(idx)>=2 else 0
tot=0
for s in seq[start:]:
    print("%-42s %10.1f us  %s"%(s[0], s[1]/1000, s[2])); tot+=s[1]
print("sum %.1f us"%(tot/1000))
PY
