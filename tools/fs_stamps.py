"""development: phase timestamps of the fused statistics kernels (SM_FS_STAMPS=1; `sample` = stop after k_fs_sample)
    python tools/fs_stamps.py [sample] [R C]"""
import sys, os
from pathlib import Path
import numpy as np
import torch
os.environ["SM_FS_STAMPS"] = "1"
args = sys.argv[1:]
only_sample = bool(args) and args[0] == "sample"
if only_sample:
    os.environ["SM_FS_ONLY_SAMPLE"] = "1"; args = args[1:]
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from shardmerge_b200 import engine as E
R, C = (int(args[0]), int(args[1])) if len(args) >= 2 else (4096, 4096)
dev = torch.device("cuda:0")
ws = E.get_workspace(R, C, dev)
ws.re[0].normal_(); ws.re[1].normal_()
N = R * C
for it in range(3):
    ws.ctl.zero_()
    E.fstats_cutoff(ws, ws.re[0], ws.re[1], int(2 * N * 0.08), 0.375)
    torch.cuda.synchronize()
total = ws.plan.lib.sm_fstats_ws_bytes(ws.plan.handle)
base = ws.sel_ws.data_ptr() + ws.sel_ws.numel() - total          # fs_carve: the last `total` bytes of the buffer
al = (base + 63) // 64 * 64 - base
off = (base + al + total - 512 - 65536 + 63) // 64 * 64 - base      # fs_carve: the stamps follow the side lists
nc = 256 if only_sample else 296
o0 = ws.sel_ws.numel() - total + off
st = ws.sel_ws[o0: o0 + nc * 8 * 8].view(torch.int64).cpu().reshape(nc, 8).numpy()
t0 = st[:, 0][st[:, 0] > 0].min()
for k in range(6):
    col = st[:, k]; col = col[col > 0]
    if len(col): print(f"stamp {k}: min {int(col.min() - t0):7d} ns  median {int(np.median(col) - t0):7d} ns  max {int(col.max() - t0):7d} ns  (n={len(col)})")
print("status", ws.fs_status())
