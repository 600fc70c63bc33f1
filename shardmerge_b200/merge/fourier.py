"""FourierMerge (earlier, in-RAM variant) with the interface of shard/merge/fourier.py:35-205 -- not what the CLI uses
(shard/__main__.py:22 imports fast_fourier), kept because it shares the hot path's tensor functions.

What differs from fast_fourier.FourierMerge, all reproduced here:
  * deltas are formed in the MODELS' OWN dtype (`ft_tensor -= base_tensor`, :120), so a bf16 model gives bf16 deltas;
  * pairs are chosen by least |mean cosine similarity| (correlate_pairs, :136; one streaming kernel per pair here:
    sm_cosine_cols) instead of by norm products; target_norm is the MEDIAN of the norms (:129);
  * the SLERP weights are looked up by POSITION in config.finetune_merge (:175-176, the reference's own TODO);
  * the arithmetic branch does not scale b (:169-170); models named in task_add_models are applied afterwards with
    task_arithmetic_fft2(agreement=False) (:193-198);
  * the result is base + merged in the promoted dtype (fp32 for bf16 models): no bf16 cast (:200-205), and a missing
    is_input / is_output model raises (:65-67, 79-81).
Everything runs on the tensors' CUDA device through shardmerge_b200.tensor.functions; there is no CPU path."""
from __future__ import annotations

import asyncio
import logging
from typing import List, Optional

import torch

from .. import engine as E
from ..config import MergeConfig
from ..constants import INPUT_LAYER, OUTPUT_LAYER
from ..tensor.functions import correlate_pairs, correlated_pairs, merge_tensors_fft2_slerp, task_arithmetic_fft2
from .base import MergeTensorsBase

logger = logging.getLogger(__name__)


def task_arithmetic(t0: torch.Tensor, t1: torch.Tensor) -> torch.Tensor:
    """t0 + t1 where the signs agree, else t0 (shard/merge/fourier.py:28-32)."""
    return torch.where(torch.sign(t0) == torch.sign(t1), t0 + t1, t0)


class FourierMerge(MergeTensorsBase):
    def __init__(self, config: MergeConfig, task_add_models: Optional[List[str]] = None,
                 target_norm_offset: float = 1e-10, cull_start_pct: float = 0.20, index_manager=None, **kwargs):
        super().__init__(config, index_manager)
        self.task_add_models = task_add_models or []
        self.target_norm_offset = target_norm_offset
        self.cull_start_pct = cull_start_pct
        self.last_info: dict = {}

    def get_readme(self) -> str:
        models = "\n".join(f"- {m.model}" for m in self.config.finetune_merge)
        return f"# SLERP-FFT Merged Model\nBase: {self.config.output_base_model}\nModels merged:\n{models}\n"

    async def _merge_layer(self, shard_layer, device: str) -> torch.Tensor:
        number, name = shard_layer.layer_number, shard_layer.layer_name
        if number in (INPUT_LAYER, OUTPUT_LAYER):
            attr, what = ("is_input", "input") if number == INPUT_LAYER else ("is_output", "output")
            chosen = next((m for m in self.config.finetune_merge if getattr(m, attr)), None)
            if chosen is None:
                raise ValueError(f"No {what} model found")                       # :65-67, :79-81
            logger.info(f"Passthrough - {name} is an {what} layer, using {chosen.model} as {what}")
            return await self.index_manager.get_tensor(chosen.model, name, device=device).get()

        dev = E._require_cuda(device)
        base = await self.index_manager.get_tensor(self.config.output_base_model, name, device=device).get()
        fts = await asyncio.gather(*[self.index_manager.get_tensor(m.model, name, device=device).get()
                                     for m in self.config.finetune_merge if m.use_layer_index(number)])
        layer_stack, add_stack, norms = [], [], []
        for i, ft in enumerate(fts):
            delta = ft - base                                                     # the models' own dtype (:120)
            model = self.config.finetune_merge[i]                                 # positional, as the reference does (:121)
            if model.model in self.task_add_models:
                add_stack.append((model.model, delta))
            else:
                norms.append(torch.norm(delta).item())
                layer_stack.append((model.model, delta))
        target_norm = torch.tensor(norms).median().item() + self.target_norm_offset   # :129
        cull_pct = self.cull_start_pct
        branches = []
        while len(layer_stack) > 1:
            correlation = correlate_pairs(torch.stack([t for _, t in layer_stack], dim=0), store_device="cpu",
                                          work_device=str(dev))
            next_stack = []
            for x, y, _ in correlated_pairs(correlation, way="least"):
                if y < 0:
                    next_stack.append(layer_stack[x])
                    continue
                (a_key, a), (b_key, b) = layer_stack[x], layer_stack[y]
                norm_a, norm_b = torch.norm(a).item(), torch.norm(b).item()
                if abs(norm_a) < abs(norm_b):
                    a, b, a_key, b_key, norm_a, norm_b = b, a, b_key, a_key, norm_b, norm_a
                cnorm_a, cnorm_b = abs(norm_a / target_norm), abs(norm_b / target_norm)
                n_ratio = cnorm_b / (cnorm_a + 1e-10)
                if cnorm_a < 1e-6:
                    merged = a + b                                                # :164-166
                    branches.append("add")
                elif cnorm_b < 1e-6 or n_ratio < 0.1:
                    scaled_a = a * target_norm / norm_a                           # :168-170 (b is not scaled here)
                    merged = task_arithmetic_fft2(scaled_a, b, t=1.0, agreement=True, device=str(dev), _result_device=dev)
                    branches.append("arith")
                else:
                    a_weight = self.config.finetune_merge[x].alpha                # positional (:175-176)
                    b_weight = self.config.finetune_merge[y].alpha
                    a_prop = a_weight / (a_weight + b_weight)
                    merged, _, _ = merge_tensors_fft2_slerp(a, b, t=a_prop, t_sum=1.0, cutoff_pct=0.08, cull_pct=cull_pct,
                                                            device=str(dev), _result_device=dev)
                    merged = merged * target_norm
                    branches.append("slerp")
                    logger.info(f"SLERP-FFT Merged {a_key} and {b_key} with weight {a_prop}")
                next_stack.append((f"{a_key}_{b_key}", merged))
            layer_stack = next_stack
            cull_pct = cull_pct / 2.0
        result = layer_stack[0][1]
        for model_name, ft in add_stack:                                          # :193-198
            result = task_arithmetic_fft2(result, ft, t=1, agreement=False, device=str(dev), _result_device=dev)
            logger.info(f"Arithmetic Merged {model_name} with weight 1")
        result = base + result.to(dev)
        if torch.any(torch.isnan(result)):
            result[torch.isnan(result)] = 0.0
        if torch.any(torch.isinf(result)):
            raise ValueError(f"Inf in merged tensor for {name}")
        self.last_info = dict(branches=branches, layer=name, target_norm=target_norm, norms=norms)
        return result
