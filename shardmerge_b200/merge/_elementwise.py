"""Shared body of AdditionMerge / TaskAdditionMerge: fetch the base and every finetune of one tensor, run ONE streaming
kernel (csrc/kernels_elem.cu through the C ABI entry sm_elem_merge) on the device, return the bf16 result.

Like the reference (shard/merge/addition.py:44-83, shard/merge/taskaddition.py:44-83) these strategies do not look at
layer numbers, alphas or layer ranges: every tensor of every finetune takes part, and what comes back is the combined
DELTA (the base is not added back)."""
from __future__ import annotations

import asyncio
import ctypes

import torch

from .. import _lib
from .. import engine as E


_DTYPE_CODE = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
MAX_MODELS = 64


def elem_merge(mode: int, base: torch.Tensor, fts, dev) -> torch.Tensor:
    """mode 0: sum of deltas; mode 1: sign-agreement sum.  CUDA tensors of one shape and one dtype (fp32 / bf16 / fp16: every
    op is rounded to that dtype, as in the reference); 1..64 finetunes."""
    code = _DTYPE_CODE.get(base.dtype)
    if code is None or any(t.dtype != base.dtype for t in fts):
        raise NotImplementedError("shardmerge_b200 element-wise strategies: base and finetunes of ONE dtype out of float32, "
                                  f"bfloat16, float16 (got {base.dtype} and {sorted({str(t.dtype) for t in fts})}); the kernels "
                                  "reproduce torch's per-op rounding, there is no CPU or mixed-dtype fallback")
    if not 1 <= len(fts) <= MAX_MODELS:
        raise NotImplementedError(f"shardmerge_b200 element-wise strategies: 1..{MAX_MODELS} finetunes, got {len(fts)}")
    if any(tuple(t.shape) != tuple(base.shape) for t in fts):
        raise ValueError("finetune / base shape mismatch")
    lib = _lib.load()
    base = base.contiguous()
    fts = [t.contiguous() for t in fts]
    out = torch.empty_like(base)
    ptrs = (ctypes.c_void_p * len(fts))(*[t.data_ptr() for t in fts])
    rc = lib.sm_elem_merge(int(mode), code, base.numel(), base.data_ptr(), ptrs, len(fts), out.data_ptr(), E._stream(dev))
    _lib.check(rc, "sm_elem_merge")
    return out


async def merge_layer_elementwise(merger, shard_layer, device: str, mode: int) -> torch.Tensor:
    dev = E._require_cuda(device)
    name = shard_layer.layer_name
    base_p = merger.index_manager.get_tensor(merger.config.output_base_model, name, device=device)
    ft_ps = [merger.index_manager.get_tensor(m.model, name, device=device).get() for m in merger.config.finetune_merge]
    base = await base_p.get()
    fts = await asyncio.gather(*ft_ps)
    return elem_merge(mode, base, list(fts), dev)
