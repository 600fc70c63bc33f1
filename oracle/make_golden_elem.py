"""make_golden_elem.py -- tests/golden/elem_*.npz: the reference's AdditionMerge / TaskAdditionMerge (shard/merge/addition.py,
shard/merge/taskaddition.py) run on CPU on seeded bf16 models, incl. zeros / inf / NaN / opposite-sign cases.
Run here: python oracle/make_golden_elem.py   (TEST INFRASTRUCTURE; the reference does not travel to the GPU box)"""
import sys, asyncio, tempfile
import numpy as np, torch
from pathlib import Path
sys.path.insert(0, "/root/reference"); sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import make_golden as MG
from shard.config import MergeConfig, MergeModel
from shard.merge.addition import AdditionMerge
from shard.merge.taskaddition import TaskAdditionMerge
from shard.writer import ShardLayer
torch.set_num_threads(4)
_SIG = (0.002, 0.0026, 0.0023, 0.0029)
_TDT = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}
def _store(t, dtype):            # bf16 travels as uint16 bit patterns, the others as numpy arrays of their own dtype
    return MG.bf16_bits(t) if dtype == "bf16" else t.numpy().copy()
def run(cls, shape, seed, n_models, special=False, dtype="bf16"):
    layer = "model.layers.3.mlp.up_proj.weight"
    base, fts = MG.synth(shape, seed, n_models, sigmas=tuple(_SIG[k % 4] * (1 + 0.07 * (k // 4)) for k in range(n_models)))
    if dtype != "bf16":          # the same models stored in another dtype (fp32 keeps extra low bits from a second draw)
        g = torch.Generator().manual_seed(seed + 7)
        base = (base.float() + (1e-5 * torch.randn(base.shape, generator=g) if dtype == "f32" else 0)).to(_TDT[dtype])
        fts = [(t.float() + (1e-5 * torch.randn(t.shape, generator=g) if dtype == "f32" else 0)).to(_TDT[dtype]) for t in fts]
    if special:          # zeros, identical models, opposite deltas, inf / nan
        fts[0].view(-1)[:8] = base.view(-1)[:8]
        fts[1].view(-1)[4:12] = base.view(-1)[4:12]
        fts[0].view(-1)[16] = float("inf"); fts[1].view(-1)[17] = float("nan"); base.view(-1)[18] = float("inf")
        fts[0].view(-1)[20:24] = (base.float().view(-1)[20:24] + 0.01).to(base.dtype)
        fts[1].view(-1)[20:24] = (base.float().view(-1)[20:24] - 0.01).to(base.dtype)
    tensors = {("org/base", layer): base}
    models = []
    for k, ft in enumerate(fts):
        tensors[(f"org/ft{k}", layer)] = ft
        models.append(MergeModel(model=f"org/ft{k}", base="org/base", alpha=1.0))
    with tempfile.TemporaryDirectory() as td:
        cfg = MergeConfig(finetune_merge=models, output_base_model="org/base", output_dir=td + "/out", cache_dir=td + "/cache", storage_dir=td + "/st")
        merger = cls(cfg, index_manager=MG._StubIndex(tensors))
        out = asyncio.run(merger._merge_layer(ShardLayer(0, "s", layer, False), "cpu"))
    assert out.dtype == _TDT[dtype], out.dtype
    d = dict(base=_store(base, dtype), out=_store(out, dtype), n=np.int64(n_models), dtype=np.array(dtype))
    for k, ft in enumerate(fts):
        d[f"ft{k}"] = _store(ft, dtype)
    return d
cases = {
    "elem_addition_3x_64x128": run(AdditionMerge, (64, 128), 31, 3),
    "elem_addition_2x_special_32x64": run(AdditionMerge, (32, 64), 32, 2, special=True),
    "elem_taskaddition_3x_64x128": run(TaskAdditionMerge, (64, 128), 33, 3),
    "elem_taskaddition_4x_96x40": run(TaskAdditionMerge, (96, 40), 34, 4),
    "elem_taskaddition_2x_special_32x64": run(TaskAdditionMerge, (32, 64), 35, 2, special=True),
    # other storage dtypes and more than 8 models (torch.sum cascades from 16 rows on)
    "elem_addition_f16_3x_64x128": run(AdditionMerge, (64, 128), 36, 3, dtype="f16"),
    "elem_addition_f32_2x_special_32x64": run(AdditionMerge, (32, 64), 37, 2, special=True, dtype="f32"),
    "elem_addition_10x_48x64": run(AdditionMerge, (48, 64), 38, 10),
    "elem_taskaddition_f16_3x_special_64x128": run(TaskAdditionMerge, (64, 128), 39, 3, special=True, dtype="f16"),
    "elem_taskaddition_f32_4x_96x40": run(TaskAdditionMerge, (96, 40), 40, 4, dtype="f32"),
    "elem_taskaddition_12x_48x64": run(TaskAdditionMerge, (48, 64), 41, 12),
    "elem_taskaddition_20x_48x64": run(TaskAdditionMerge, (48, 64), 42, 20),
    "elem_taskaddition_f32_35x_24x64": run(TaskAdditionMerge, (24, 64), 43, 35, dtype="f32"),
}
for name, d in cases.items():
    np.savez_compressed(MG.OUT / f"{name}.npz", **d)
print("wrote", list(cases))
# check the oracle right away
from oracle import oracle_np as O
for name, d in cases.items():
    fts = [d[f"ft{k}"] for k in range(int(d["n"]))]
    dt = str(d["dtype"])
    got = (O.addition_merge if "task" not in name else O.taskaddition_merge)(d["base"], fts, dt)
    wide = (lambda a: O.bf16_to_f32(a)) if dt == "bf16" else (lambda a: np.asarray(a, dtype=np.float32))
    same = (got == d["out"]) | ((wide(got) != wide(got)) & (wide(d["out"]) != wide(d["out"])))
    print(name, "bit-exact (NaN == NaN):", bool(same.all()), "mismatches", int((~same).sum()))
