"""Shared body of AdditionMerge / TaskAdditionMerge: fetch the base and every finetune of one tensor, run ONE streaming
kernel (csrc/kernels_elem.cu through the C ABI entry sm_elem_merge_bf16) on the device, return the bf16 result.

Like the reference (shard/merge/addition.py:44-83, shard/merge/taskaddition.py:44-83) these strategies do not look at
layer numbers, alphas or layer ranges: every tensor of every finetune takes part, and what comes back is the combined
DELTA (the base is not added back)."""
from __future__ import annotations

import asyncio
import ctypes

import torch

from .. import _lib
from .. import engine as E


def elem_merge(mode: int, base: torch.Tensor, fts, dev) -> torch.Tensor:
    """mode 0: sum of deltas; mode 1: sign-agreement sum.  bf16 CUDA tensors of one shape; at most 8 finetunes."""
    if base.dtype != torch.bfloat16 or any(t.dtype != torch.bfloat16 for t in fts):
        raise NotImplementedError("shardmerge_b200 element-wise strategies: bfloat16 models only (the kernels reproduce "
                                  "torch's per-op bf16 rounding; there is no CPU or other-dtype fallback)")
    if any(tuple(t.shape) != tuple(base.shape) for t in fts):
        raise ValueError("finetune / base shape mismatch")
    lib = _lib.load()
    base = base.contiguous()
    fts = [t.contiguous() for t in fts]
    out = torch.empty_like(base)
    ptrs = (ctypes.c_void_p * len(fts))(*[t.data_ptr() for t in fts])
    rc = lib.sm_elem_merge_bf16(int(mode), base.numel(), base.data_ptr(), ptrs, len(fts), out.data_ptr(), E._stream(dev))
    _lib.check(rc, "sm_elem_merge_bf16")
    return out


async def merge_layer_elementwise(merger, shard_layer, device: str, mode: int) -> torch.Tensor:
    dev = E._require_cuda(device)
    name = shard_layer.layer_name
    base_p = merger.index_manager.get_tensor(merger.config.output_base_model, name, device=device)
    ft_ps = [merger.index_manager.get_tensor(m.model, name, device=device).get() for m in merger.config.finetune_merge]
    base = await base_p.get()
    fts = await asyncio.gather(*ft_ps)
    return elem_merge(mode, base, list(fts), dev)
