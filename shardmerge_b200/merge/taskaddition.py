"""TaskAdditionMerge -- same class and `_merge_layer(shard_layer, device) -> Tensor` as shard/merge/taskaddition.py:27-83:
per element, the deltas whose sign agrees with the majority sign over the finetunes are summed, the others dropped.
One streaming sm_100a kernel per tensor."""
from __future__ import annotations

import logging

import torch

from ..writer import ShardLayer
from ._elementwise import merge_layer_elementwise
from .base import MergeTensorsBase

logger = logging.getLogger(__name__)


class TaskAdditionMerge(MergeTensorsBase):
    """Addition merge operation, using sign agreement"""

    def get_readme(self) -> str:
        return f"""# Merged Model

Base Model: {self.config.output_base_model}
Finetuned Models:
{chr(10).join('- ' + model.model for model in self.config.finetune_merge)}

This model was created by computing and combining the delta weights
from each finetuned model relative to the base model, using sign agreement.
"""

    async def _merge_layer(self, shard_layer: ShardLayer, device: str) -> torch.Tensor:
        logger.info(f"Processing layer: {shard_layer.layer_name}")
        return await merge_layer_elementwise(self, shard_layer, device, mode=1)
