#!/bin/bash
# quick GPU check used during development: parity tests, single-tensor timings, short bench with per-class profile
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 400 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
for shp in "14336 4096" "4096 4096" "1024 4096" "4096 14336" "1 4096"; do python tools/profile_one.py $shp 5 | tail -1; done
timeout 600 python bench.py --layers ${LAYERS:-4} --steps 3 --warmup 3 --no-cpu-baseline ${E2E:---no-e2e} --profile-json gpurun_out/bench_profile.json > gpurun_out/bench.log 2>&1
python - <<PY
import json
d=json.load(open("gpurun_out/bench_profile.json"))
print("value %.2f Gparam/s  ms/step %.2f  e2e %s"%(d["line"]["value"]/1e9, d["line"]["ms_per_step"], d["line"].get("e2e")))
tot=0
for k,v in d["summary"].items():
    print("%-12s calls %3d launches %4d ms %8.3f  GB/s %7.1f"%(k,v["calls"],v["launches"],v["ms"] or 0,(v["bytes"]/v["ms"]/1e6) if v["ms"] else 0)); tot+=v["ms"] or 0
print("kernel ms total %.2f steps ms %.2f"%(tot, d["line"]["ms_per_step"]*d["line"]["steps"]))
PY
