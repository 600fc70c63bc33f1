"""Profile driver: merge ONE synthetic tensor with M finetunes (pair tree) a few times:  python tools/profile_tree.py R C M [iters]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from shardmerge_b200 import engine as E
from shardmerge_b200.config import MergeConfig
from shardmerge_b200.index import InMemoryIndex
from shardmerge_b200.merge.fast_fourier import FourierMerge
R, C, M = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
base = (0.02 * torch.randn((R, C), generator=g, device=dev)).to(torch.bfloat16)
sig, al = (0.002, 0.0026, 0.0023, 0.0029), (0.3, 0.5, 0.4, 0.2)
fts = [(base.float() + s * torch.randn((R, C), generator=g, device=dev)).to(torch.bfloat16) for s in sig[:M]]
fm = FourierMerge(MergeConfig(finetune_merge=[], output_base_model="b", output_dir="/tmp/unused"), index_manager=InMemoryIndex({}))
for it in range(iters):
    srcs = [E.make_source(base, ft, weight=a, name=f"m{k}") for k, (ft, a) in enumerate(zip(fts, al))]
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fm.merge_sources(srcs, base, dev, layer_name="model.layers.0.profile")
    e1.record(); torch.cuda.synchronize()
    print(f"iter {it}: {e0.elapsed_time(e1):.3f} ms  branches={fm.last_info['branches']}")
