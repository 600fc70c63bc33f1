"""AdditionMerge -- same class, constructor and `_merge_layer(shard_layer, device) -> Tensor` as
shard/merge/addition.py:27-83: the sum over all finetunes of (finetune - base), computed in the tensors' own dtype
(note: the base is not added back; that is what the reference returns).  One streaming sm_100a kernel per tensor."""
from __future__ import annotations

import logging

import torch

from ..writer import ShardLayer
from ._elementwise import merge_layer_elementwise
from .base import MergeTensorsBase

logger = logging.getLogger(__name__)


class AdditionMerge(MergeTensorsBase):
    """Simple addition merge operation"""

    def get_readme(self) -> str:
        return f"""# Merged Model

Base Model: {self.config.output_base_model}
Finetuned Models:
{chr(10).join('- ' + model.model for model in self.config.finetune_merge)}

This model was created by computing and combining the delta weights
from each finetuned model relative to the base model.
"""

    async def _merge_layer(self, shard_layer: ShardLayer, device: str = "cuda") -> torch.Tensor:
        logger.info(f"Processing layer: {shard_layer.layer_name}")
        return await merge_layer_elementwise(self, shard_layer, device, mode=0)
