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
hc = ws.sel_ws[al: al + kBins * 8].view(torch.int64).cpu()
scnt = ws.sel_ws[al + kBins * 8 + kBins * 24: al + kBins * 8 + kBins * 24 + kBins * 4].view(torch.int32).cpu()
print("bucket entries total", int((hc & 0xffffffff).sum()), "max", int((hc & 0xffffffff).max()), "keys", int((hc >> 32).sum()), "bcap", bcap)
print("side entries total", int(scnt.sum()), "max", int(scnt.max()), "scap", scap)
h = ws.ctl.cpu()
o = 512
import struct
print("state lo hi shift status", struct.unpack_from("<QQIIII", bytes(h[o:o+32].tolist())))
d1 = (st[:, 1] - st[:, 0]).float(); d2 = (st[:, 2] - st[:, 1]).float()
import numpy as np
for nm, d in (("loop", d1), ("flush", d2)):
    q = np.percentile(d.numpy(), [0, 10, 50, 90, 99, 100])
    print(nm, "percentiles ns", [int(v) for v in q])
tot = (st[:, 2] - st[:, 0])
idx = torch.argsort(tot, descending=True)[:12]
print("slowest CTAs (block, smid, loop ns, flush ns):", [(int(i), int(st[i, 6]), int(d1[i]), int(d2[i])) for i in idx])
