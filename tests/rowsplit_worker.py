"""Worker of tests/test_gpu_rowsplit.py (TEST INFRASTRUCTURE): one process per GPU under torchrun.
    python -m torch.distributed.run --nproc-per-node G tests/rowsplit_worker.py R C [reference]
Every rank generates the same full tensors from a seed, takes its block of rows, runs shardmerge_b200.rowsplit.merge_rowsplit;
rank 0 gathers the blocks and compares them with the one-GPU FourierMerge.merge_sources on the full tensor (and, with
`reference`, with the unmodified reference on device="cuda", oracle/ref_runner.py)."""
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    R, C = int(sys.argv[1]), int(sys.argv[2])
    with_ref = len(sys.argv) > 3 and sys.argv[3] == "reference"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from shardmerge_b200 import engine as E
    from shardmerge_b200.rowsplit import merge_rowsplit
    from tests.parity_util import bf16_ulp_distance

    assert R % world == 0
    Rl = R // world
    g = torch.Generator(device=dev).manual_seed(4242)
    # the same bits on every rank (same generator state), generated block by block to bound memory on big shapes
    def full(scale, offset=None):
        return (scale * torch.randn((R, C), generator=g, device=dev))
    base = full(0.02).to(torch.bfloat16)
    fts = [(base.float() + s * torch.randn((R, C), generator=g, device=dev)).to(torch.bfloat16) for s in (0.002, 0.0026)]
    rows = slice(rank * Rl, (rank + 1) * Rl)
    info = {}
    out = merge_rowsplit(base[rows].clone(), [f[rows].clone() for f in fts], (0.3, 0.5), info=info)      # warm-up: plans, NCCL
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    out = merge_rowsplit(base[rows].clone(), [f[rows].clone() for f in fts], (0.3, 0.5), info=info)
    torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t0
    blocks = [torch.empty_like(out) for _ in range(world)] if rank == 0 else None
    dist.gather(out, blocks, dst=0)
    if rank == 0:
        got = torch.cat(blocks, 0)
        bits = lambda t: t.contiguous().view(torch.int16).cpu().numpy().view(np.uint16)
        res = dict(R=R, C=C, ranks=world, seconds=round(dt, 4), params_per_s=R * C / dt, info=info)
        del blocks
        E.clear_caches(); torch.cuda.empty_cache()
        from shardmerge_b200.config import MergeConfig
        from shardmerge_b200.index import InMemoryIndex
        from shardmerge_b200.merge.fast_fourier import FourierMerge
        fm = FourierMerge(MergeConfig(finetune_merge=[], output_base_model="b", output_dir="/tmp/unused"), index_manager=InMemoryIndex({}))
        one = fm.merge_sources([E.make_source(base, f, weight=a, name=f"m{k}") for k, (f, a) in enumerate(zip(fts, (0.3, 0.5)))],
                               base, dev, layer_name="model.layers.0.x")
        u = bf16_ulp_distance(bits(got), bits(one))
        res.update(vs_one_gpu_exact=float((u == 0).mean()), vs_one_gpu_within_1ulp=float((u <= 1).mean()),
                   one_gpu_target_norm=fm.last_info["target_norm"])
        if with_ref:
            from oracle import ref_runner as RR
            if RR.available():
                del one
                E.clear_caches(); torch.cuda.empty_cache()
                try:
                    t1 = time.perf_counter()
                    ref = RR.merge_layer(base, fts, [0.3, 0.5], device=str(dev))
                    torch.cuda.synchronize()
                    ur = bf16_ulp_distance(bits(got), bits(ref))
                    res.update(vs_reference_cuda_exact=float((ur == 0).mean()), vs_reference_cuda_within_1ulp=float((ur <= 1).mean()),
                               reference_cuda_seconds=round(time.perf_counter() - t1, 2))
                except Exception as exc:                     # e.g. the reference's sorts / temporaries do not fit at 1 G elements
                    res.update(reference_cuda_error=f"{type(exc).__name__}: {str(exc)[:200]}")
        print("ROWSPLIT " + json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
