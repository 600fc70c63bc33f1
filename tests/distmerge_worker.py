"""Worker of tests/test_gpu_rowsplit.py::test_merge_distributed_two_gpus (TEST INFRASTRUCTURE): one process per GPU under torchrun.
    python -m torch.distributed.run --nproc-per-node G tests/distmerge_worker.py OUT_DIR
Every rank builds the same toy model (seeded), runs shardmerge_b200.schedule.merge_distributed with the real FourierMerge on its
own GPU -- whole output shards per rank, no collective on the data path -- and rank 0 verifies the output directory."""
import asyncio
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def toy_models(L=6):
    g = torch.Generator().manual_seed(11)
    names = {"model.embed_tokens.weight": (64, 512), "model.norm.weight": (2048,), "lm_head.weight": (64, 512)}
    for l in range(L):
        names[f"model.layers.{l}.mlp.up_proj.weight"] = (1024, 2048)
        names[f"model.layers.{l}.self_attn.k_proj.weight"] = (256, 2048)
        names[f"model.layers.{l}.input_layernorm.weight"] = (2048,)
    models = {"org/base": {}, "org/ft0": {}, "org/ft1": {}}
    for n, shape in names.items():
        one_d = len(shape) == 1
        base = ((1.0 if one_d else 0.0) + (0.1 if one_d else 0.02) * torch.randn(shape, generator=g)).to(torch.bfloat16)
        models["org/base"][n] = base
        for k, s in enumerate((0.002, 0.0026)):
            models[f"org/ft{k}"][n] = (base.float() + (5 * s if one_d else s) * torch.randn(shape, generator=g)).to(torch.bfloat16)
    return models


def main():
    out_dir = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    from shardmerge_b200 import schedule
    from shardmerge_b200.config import MergeConfig, MergeModel
    from shardmerge_b200.index import InMemoryIndex
    from shardmerge_b200.merge.fast_fourier import FourierMerge
    from shardmerge_b200.validate import verify_output
    models = toy_models()
    cfg = MergeConfig(finetune_merge=[MergeModel(model="org/ft0", base="org/base", alpha=0.3, is_input=True),
                                      MergeModel(model="org/ft1", base="org/base", alpha=0.5, is_output=True)],
                      output_base_model="org/base", output_dir=out_dir, device=dev)
    fm = FourierMerge(cfg, index_manager=InMemoryIndex(models))
    mine = asyncio.run(schedule.merge_distributed(fm, dev, rank, world, barrier=dist.barrier))
    dist.barrier()
    res = dict(rank=rank, shards=mine)
    if rank == 0:
        rep = verify_output(out_dir, expected_dtype=torch.bfloat16, expected_shapes={n: tuple(t.shape) for n, t in models["org/base"].items()})
        res.update(ok=rep.ok, summary=rep.summary(), tensors=rep.tensors, readme=(Path(out_dir) / "README.md").exists())
    print("DISTMERGE " + json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
