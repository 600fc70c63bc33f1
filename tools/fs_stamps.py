"""development: phase timestamps of k_fs_sample (SM_FS_STAMPS=1 SM_FS_ONLY_SAMPLE=1)"""
import sys, os
from pathlib import Path
import torch
os.environ["SM_FS_STAMPS"] = "1"
if len(sys.argv) > 1 and sys.argv[1] == "sample": os.environ["SM_FS_ONLY_SAMPLE"] = "1"
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from shardmerge_b200 import engine as E
R, C = 4096, 4096
dev = torch.device("cuda:0")
ws = E.get_workspace(R, C, dev)
ws.re[0].normal_(); ws.re[1].normal_()
N = R * C
for it in range(3):
    ws.ctl.zero_()
    E.fstats_cutoff(ws, ws.re[0], ws.re[1], int(2 * N * 0.08), 0.375)
    torch.cuda.synchronize()
# locate dbg = bkt pointer: scan the workspace tail is awkward; recompute the carve offsets like fs_carve
import ctypes
kBins = 2048
lib = ws.plan.lib
total = lib.sm_fstats_ws_bytes(ws.plan.handle)
base = ws.sel_ws.data_ptr(); al = (base + 63) // 64 * 64 - base
# scap/bcap as in fs_bcap / fs_scap
import math
frac = 2.0 * (6.0 * math.sqrt(65536 * 0.25) + 16.0) / 65536
Ch = C // 2
scap = int(frac * R * (Ch + 1) / 1024.0 * 4.0) + 64
bcap = int(frac * R * (Ch + 1) * 2.0 / 1024.0 * 4.0) + 64
off = al + kBins * 8 + kBins * 24 + kBins * 4 + (1 << 13) * 4 + kBins * scap * 16 + kBins * bcap * 4
off = (base + off + 63) // 64 * 64 - base
nc = 592 if 'SM_FS_ONLY_SAMPLE' not in os.environ else 64
st = ws.sel_ws[off: off + nc * 8 * 8].view(torch.int64).cpu().reshape(nc, 8)
t0 = int(st[:, 0].min())
for k in range(6):
    col = st[:, k]; col = col[col > 0]
    if len(col): print(f"stamp {k}: min {int(col.min()) - t0:7d} ns  max {int(col.max()) - t0:7d} ns  (n={len(col)})")
