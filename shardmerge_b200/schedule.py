"""Work partition of a model merge across the GPUs of one box.

The reference merges tensors in one sequential loop on one device (shard/merge/base.py:186-196,
"This could be done in parallel, but we would need to use multiprocessing").  Every
`_merge_layer` call depends only on that tensor's slices of the models
(shard/merge/fast_fourier.py:103-276), so the work shards by whole tensors with no collective
on the data path.

Two granularities:
  * tensor_partition  -- longest-processing-time-first over (tensor, numel): what `bench.py --gpus N` uses to
                         split ONE model's tensor list over the ranks (strong scaling, configs 3/4).
  * shard_partition   -- LPT over whole output shards: every rank owns complete safetensors files,
                         so the writer needs no cross-rank assembly (what merge_distributed uses).
One process per GPU; torch.distributed is used only for rendezvous / barriers.
"""
from __future__ import annotations

import heapq
from typing import Dict, List, Sequence, Tuple


def lpt(items: Sequence[Tuple[str, int]], n_ranks: int) -> List[List[str]]:
    """Longest-processing-time-first: heaviest item to the currently lightest rank.
    Deterministic (ties broken by name then rank), so every rank computes the same answer."""
    if n_ranks < 1:
        raise ValueError("n_ranks must be >= 1")
    order = sorted(items, key=lambda it: (-int(it[1]), it[0]))
    heap = [(0, r) for r in range(n_ranks)]
    heapq.heapify(heap)
    out: List[List[str]] = [[] for _ in range(n_ranks)]
    for name, cost in order:
        load, r = heapq.heappop(heap)
        out[r].append(name)
        heapq.heappush(heap, (load + int(cost), r))
    return out


def merge_cost(numel: int, n_models: int) -> int:
    """Relative cost of one tensor: (n_models - 1) pair merges, each ~ one forward transform per
    input plus statistics and one inverse transform (DESIGN.md byte model)."""
    pairs = max(n_models - 1, 0)
    return int(numel) * max(1, 24 * n_models + 36 * pairs)


def tensor_partition(tensor_numels: Dict[str, int], n_ranks: int, n_models: int = 2) -> List[List[str]]:
    return lpt([(n, merge_cost(v, n_models)) for n, v in tensor_numels.items()], n_ranks)


def shard_partition(weight_map: Dict[str, str], tensor_numels: Dict[str, int], n_ranks: int,
                    n_models: int = 2) -> List[List[str]]:
    """-> per rank, the list of shard file names it owns."""
    cost: Dict[str, int] = {}
    for tensor, shard in weight_map.items():
        cost[shard] = cost.get(shard, 0) + merge_cost(tensor_numels.get(tensor, 0), n_models)
    return lpt(list(cost.items()), n_ranks)


def imbalance(parts: List[List[str]], cost: Dict[str, int]) -> float:
    loads = [sum(cost[n] for n in p) for p in parts]
    mean = sum(loads) / len(loads)
    return (max(loads) / mean - 1.0) if mean > 0 else 0.0


async def merge_distributed(merger, device: str, rank: int, world: int, barrier=None):
    """Shard-granular multi-GPU merge: rank r merges and writes the shards LPT assigns to it; rank 0
    also writes README.md; a barrier closes the job.  `merger` is a MergeTensorsBase."""
    await merger.initialize()
    im = merger.index_manager
    base = merger.config.output_base_model
    layer_order = im.get_layer_order(base)
    weight_map = merger.index_doc["weight_map"]
    numels = getattr(im, "tensor_numels", lambda m: {n: 1 for n in weight_map})(base)
    n_models = len(merger.config.finetune_merge)
    mine = set(shard_partition(weight_map, numels, world, n_models)[rank])
    # rank 0 alone creates the output directory's index copy (tmp file + rename); the others open it afterwards.
    # Every rank scans and writes only its own shards, so no rank ever opens a file another one is writing.
    if rank == 0:
        writer = merger.get_writer(layer_order, only_shards=mine, write_index=True)
    if barrier is not None:
        barrier()
    if rank != 0:
        writer = merger.get_writer(layer_order, only_shards=mine, write_index=(barrier is None))
    merger.defer_checks = getattr(merger, "pipeline_depth", 0) > 0
    try:
        for group in writer.shard_layers():
            if not group or group[0].shard_name not in mine:
                continue
            await merger._process_layers(writer, [sl for sl in group if not sl.written], device)
    except BaseException:
        writer.flush_partial()
        raise
    finally:
        merger.defer_checks = False
    writer.finalize(only_shards=mine)
    if barrier is not None:
        barrier()
    if rank == 0:
        with open(merger.config.output_path / "README.md", "w") as fh:
            fh.write(merger.get_readme() or "No README defined")
    return sorted(mine)
