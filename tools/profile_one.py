"""Profile driver: merge ONE synthetic tensor a few times (for ncu / compute-sanitizer runs).
    python tools/profile_one.py R C [iters]
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from shardmerge_b200 import engine as E
from shardmerge_b200.config import MergeConfig
from shardmerge_b200.index import InMemoryIndex
from shardmerge_b200.merge.fast_fourier import FourierMerge

R, C = int(sys.argv[1]), int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
shape = (R, C) if R > 1 else (C,)
base = (0.02 * torch.randn(shape, generator=g, device=dev)).to(torch.bfloat16)
fts = [(base.float() + s * torch.randn(shape, generator=g, device=dev)).to(torch.bfloat16) for s in (0.002, 0.0026)]
fm = FourierMerge(MergeConfig(finetune_merge=[], output_base_model="b", output_dir="/tmp/unused"), index_manager=InMemoryIndex({}))
for it in range(iters):
    srcs = [E.make_source(base, ft, weight=a, name=f"m{k}") for k, (ft, a) in enumerate(zip(fts, (0.3, 0.5)))]
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fm.merge_sources(srcs, base, dev, layer_name="model.layers.0.profile")
    e1.record(); torch.cuda.synchronize()
    print(f"iter {it}: {e0.elapsed_time(e1):.3f} ms  {R * C / e0.elapsed_time(e1) / 1e6:.2f} Gparam/s  branches={fm.last_info['branches']}")
