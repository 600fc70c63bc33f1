import sys, json, subprocess, os
# experiment: cost of the 1-D tensors in the timed region (monkeypatch layer_tensors)
from pathlib import Path; ROOT = str(Path(__file__).resolve().parent.parent); sys.path.insert(0, ROOT); os.chdir(ROOT)
import bench
orig = bench.layer_tensors
for mode in ("all", "no1d"):
    bench.layer_tensors = (lambda a: [t for t in orig(a) if len(t[1]) == 2]) if mode == "no1d" else orig
    sys.argv = ["bench.py", "--layers", "8", "--steps", "3", "--warmup", "3", "--no-cpu-baseline", "--no-e2e"]
    import io, contextlib
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        bench.main()
    d = json.loads(buf.getvalue().strip().splitlines()[-1])
    print(mode, "value %.2f G ms/step %.3f params %d" % (d["value"] / 1e9, d["ms_per_step"], d["config"]["merged_params_per_step"]))
