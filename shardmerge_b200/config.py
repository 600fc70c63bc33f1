"""YAML merge configuration -- schema-compatible with shard/config.py:25-126 so existing
config files load unchanged.  Only the schema is mirrored (the reference's config module is
out of the hot path); field names, defaults and required keys are identical."""
from __future__ import annotations

from dataclasses import dataclass, fields
from pathlib import Path
from typing import List

import click
import torch
import yaml

_REQUIRED = ("output_base_model", "finetune_merge", "output_dir")


@dataclass
class MergeModel:
    model: str
    base: str
    alpha: float = 1.0
    is_input: bool = False
    is_output: bool = False
    is_norm: bool = False
    start_layer: int = 0
    end_layer: int = -1

    def use_layer_index(self, layer_index: int) -> bool:
        """Layer range filter (shard/config.py:35-40): [start_layer, end_layer], end -1 = open."""
        if layer_index < self.start_layer:
            return False
        return self.end_layer == -1 or layer_index <= self.end_layer


@dataclass
class MergeConfig:
    finetune_merge: List[MergeModel]
    output_base_model: str
    output_dir: str
    output_dtype: str = "bfloat16"
    device: str = "cpu"
    clean_cache: bool = False
    cache_dir: str = "cache"
    storage_dir: str = "storage"

    @property
    def input_model(self):
        return next((m for m in self.finetune_merge if m.is_input), None)

    @property
    def output_model(self):
        return next((m for m in self.finetune_merge if m.is_output), None)

    @property
    def output_path(self) -> Path:
        return Path(self.output_dir)

    @property
    def cache_path(self) -> Path:
        return Path(self.cache_dir)

    @property
    def storage_path(self) -> Path:
        return Path(self.storage_dir)

    @property
    def output_astype(self) -> torch.dtype:
        return getattr(torch, self.output_dtype)

    def update(self, config: dict | None = None, **kwargs):
        known = {f.name for f in fields(self)}
        for src in (config or {}), kwargs:
            for key, value in src.items():
                if key in known or hasattr(self, key):
                    setattr(self, key, value)

    def to_dict(self) -> dict:
        return dict(output_base_model=self.output_base_model,
                    finetune_merge=[m.model for m in self.finetune_merge],
                    output_dir=self.output_dir, device=self.device, clean_cache=self.clean_cache,
                    cache_dir=self.cache_dir, storage_dir=self.storage_dir)

    @classmethod
    def from_yaml(cls, config_path) -> "MergeConfig":
        with open(config_path) as fh:
            raw = yaml.safe_load(fh)
        missing = [k for k in _REQUIRED if k not in raw]
        if missing:
            raise click.BadParameter(f"Missing required configuration fields: {', '.join(missing)}")   # shard/config.py:110-115
        if not isinstance(raw["finetune_merge"], list):
            raise click.BadParameter("finetune_merge must be a list of model URIs")                    # :118-121
        raw["finetune_merge"] = [MergeModel(**entry) for entry in raw["finetune_merge"]]
        return cls(**raw)
