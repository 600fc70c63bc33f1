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
def run(cls, shape, seed, n_models, special=False):
    layer = "model.layers.3.mlp.up_proj.weight"
    base, fts = MG.synth(shape, seed, n_models)
    if special:          # zeros, identical models, opposite deltas, inf / nan
        fts[0].view(-1)[:8] = base.view(-1)[:8]
        fts[1].view(-1)[4:12] = base.view(-1)[4:12]
        fts[0].view(-1)[16] = float("inf"); fts[1].view(-1)[17] = float("nan"); base.view(-1)[18] = float("inf")
        fts[0].view(-1)[20:24] = (base.float().view(-1)[20:24] + 0.01).to(torch.bfloat16)
        fts[1].view(-1)[20:24] = (base.float().view(-1)[20:24] - 0.01).to(torch.bfloat16)
    tensors = {("org/base", layer): base}
    models = []
    for k, ft in enumerate(fts):
        tensors[(f"org/ft{k}", layer)] = ft
        models.append(MergeModel(model=f"org/ft{k}", base="org/base", alpha=1.0))
    with tempfile.TemporaryDirectory() as td:
        cfg = MergeConfig(finetune_merge=models, output_base_model="org/base", output_dir=td + "/out", cache_dir=td + "/cache", storage_dir=td + "/st")
        merger = cls(cfg, index_manager=MG._StubIndex(tensors))
        out = asyncio.run(merger._merge_layer(ShardLayer(0, "s", layer, False), "cpu"))
    assert out.dtype == torch.bfloat16, out.dtype
    d = dict(base=MG.bf16_bits(base), out=MG.bf16_bits(out), n=np.int64(n_models))
    for k, ft in enumerate(fts):
        d[f"ft{k}"] = MG.bf16_bits(ft)
    return d
cases = {
    "elem_addition_3x_64x128": run(AdditionMerge, (64, 128), 31, 3),
    "elem_addition_2x_special_32x64": run(AdditionMerge, (32, 64), 32, 2, special=True),
    "elem_taskaddition_3x_64x128": run(TaskAdditionMerge, (64, 128), 33, 3),
    "elem_taskaddition_4x_96x40": run(TaskAdditionMerge, (96, 40), 34, 4),
    "elem_taskaddition_2x_special_32x64": run(TaskAdditionMerge, (32, 64), 35, 2, special=True),
}
for name, d in cases.items():
    np.savez_compressed(MG.OUT / f"{name}.npz", **d)
print("wrote", list(cases))
# check the oracle right away
from oracle import oracle_np as O
for name, d in cases.items():
    fts = [d[f"ft{k}"] for k in range(int(d["n"]))]
    got = O.addition_merge(d["base"], fts) if "addition_" in name and "task" not in name else O.taskaddition_merge(d["base"], fts)
    same = (got == d["out"]) | ((O.bf16_to_f32(got) != O.bf16_to_f32(got)) & (O.bf16_to_f32(d["out"]) != O.bf16_to_f32(d["out"])))
    print(name, "bit-exact (NaN == NaN):", bool(same.all()), "mismatches", int((~same).sum()))
