"""ModelWriter / ShardLayer with the interface of shard/writer.py.

Observable contract kept from the reference: the output directory holds a copy of the base
model's model.safetensors.index.json (:75-81), one safetensors file per entry of its weight_map
with metadata {"format": "pt"} (:143), tensors cast to output_astype (:133); existing shards are
scanned at start and their tensors count as written (resume, :93-113); finalize() raises if a
tensor is missing (:151-161).

What changes is the I/O schedule: the reference re-reads and re-writes the
whole shard for every tensor (:125-143, O(k^2) bytes per shard).  Here a shard's tensors are
staged in pinned host buffers by asynchronous device-to-host copies on a side stream -- the
merge of the next tensor overlaps the copy -- and each shard file is written exactly once, when
its last tensor has arrived (or at finalize() for shards completed across a resumed run).
"""
from __future__ import annotations

import json
import logging
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, Generator, List, Set

import torch
from safetensors import safe_open
from safetensors.torch import save_file

from .constants import INPUT_LAYER, OUTPUT_LAYER

logger = logging.getLogger(__name__)


@dataclass
class ShardLayer:
    layer_order_idx: int
    shard_name: str
    layer_name: str
    written: bool

    @property
    def layer_number(self) -> int:
        """embed -> INPUT_LAYER, final norm / lm_head -> OUTPUT_LAYER, model.layers.<n>.* -> n
        (shard/writer.py:39-57)."""
        name = self.layer_name
        if name.startswith("model.embed_tokens.weight"):
            return INPUT_LAYER
        if name.startswith(("model.norm.weight", "lm_head.weight")):
            return OUTPUT_LAYER
        if name.startswith("model.layers."):
            field_ = name.split(".")[2]
            number = int(field_)
            if str(number) == field_:
                return number
        raise ValueError(f"Unknown layer name: {name}")


@dataclass
class ModelWriter:
    base_index: dict
    output_path: Path
    layer_order: list
    output_astype: torch.dtype
    written_shard_layers: Set[tuple] = field(default_factory=set)
    shard_to_tensors: Dict[str, Set[str]] = field(default_factory=dict)

    def __post_init__(self):
        self.output_path = Path(self.output_path)
        self.output_path.mkdir(parents=True, exist_ok=True)
        self.index_path = self.output_path / "model.safetensors.index.json"
        if self.index_path.exists():
            logger.info(f"Index already exists: {self.index_path}")
            with open(self.index_path) as fh:
                self.base_index = json.load(fh)
        else:
            with open(self.index_path, "w") as fh:
                json.dump(self.base_index, fh, indent=2)
        self.shard_to_tensors = {}
        for tensor_name, shard_name in self.base_index["weight_map"].items():
            self.shard_to_tensors.setdefault(shard_name, set()).add(tensor_name)
        self._order_pos = {n: i for i, n in enumerate(self.layer_order)}
        self._staged: Dict[str, Dict[str, torch.Tensor]] = {}      # shard -> {tensor: pinned host tensor}
        self._events: Dict[str, list] = {}
        self._copy_stream = None
        self._check_existing_shards()

    # ------------------------------------------------------------------ resume
    def _check_existing_shards(self):
        for shard_name, names in self.shard_to_tensors.items():
            path = self.output_path / shard_name
            if not path.exists():
                continue
            try:
                with safe_open(path, framework="pt") as f:
                    for key in f.keys():
                        if key not in names:
                            raise ValueError(f"Tensor {key} found in {path} but not in base model")
                        self.written_shard_layers.add((shard_name, key))
            except Exception as exc:
                logger.error(f"Error validating shard {shard_name}: {exc}")
                raise

    # ------------------------------------------------------------------ staging
    def _stage(self, shard_name: str, layer_name: str, tensor: torch.Tensor):
        t = tensor.detach()
        if t.device.type == "cuda":
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=t.device)
            src = t if t.dtype == self.output_astype else t.to(self.output_astype)
            host = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
            produced = torch.cuda.current_stream(t.device).record_event()
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(produced)
                host.copy_(src, non_blocking=True)
                src.record_stream(self._copy_stream)
                done = self._copy_stream.record_event()
            self._events.setdefault(shard_name, []).append(done)
        else:
            host = t.clone().to(self.output_astype)
        self._staged.setdefault(shard_name, {})[layer_name] = host

    def _flush(self, shard_name: str):
        staged = self._staged.get(shard_name)
        if not staged:
            return
        for ev in self._events.pop(shard_name, []):
            ev.synchronize()
        path = self.output_path / shard_name
        tensors = dict(staged)
        if path.exists():                                   # resumed run: keep what an earlier run wrote
            with safe_open(path, framework="pt") as f:
                for key in f.keys():
                    tensors.setdefault(key, f.get_tensor(key))
        ordered = {n: tensors[n] for n in sorted(tensors, key=lambda n: self._order_pos.get(n, 1 << 30))}
        try:
            save_file(ordered, str(path), metadata={"format": "pt"})
            for name in staged:
                self.written_shard_layers.add((shard_name, name))
            logger.info(f"Wrote shard {shard_name} ({len(ordered)} tensors)")
        except Exception as exc:                             # reference behaviour: log, drop the file
            logger.error(f"Error saving shard {shard_name}: {exc}")
            if path.exists():
                path.unlink()
        self._staged.pop(shard_name, None)

    # ------------------------------------------------------------------ reference API
    def add_tensor(self, layer_name: str, tensor: torch.Tensor):
        shard_name = self.base_index["weight_map"][layer_name]
        if (shard_name, layer_name) in self.written_shard_layers:
            logger.info(f"Skipping {layer_name} as it's already in written shard {shard_name}")
            return
        self._stage(shard_name, layer_name, tensor)
        have = {n for (s, n) in self.written_shard_layers if s == shard_name} | set(self._staged[shard_name])
        if have >= self.shard_to_tensors[shard_name]:
            self._flush(shard_name)

    def finalize(self, only_shards=None):
        """Flush what is staged and verify completeness (shard/writer.py:151-161).  `only_shards`
        restricts the check to the shards this process owns (multi-GPU merge, schedule.py)."""
        for shard_name in list(self._staged):
            self._flush(shard_name)
        missing = [(s, n) for s, names in self.shard_to_tensors.items() for n in names
                   if (only_shards is None or s in only_shards) and (s, n) not in self.written_shard_layers]
        if missing:
            logger.error(f"Failed to write all layers. Missing: {missing}")
            raise RuntimeError(f"Incomplete model output: missing {len(missing)} layers")

    def shard_layers(self) -> Generator[List[ShardLayer], None, None]:
        for shard_name in sorted(self.shard_to_tensors):
            names = sorted(self.shard_to_tensors[shard_name], key=lambda n: self.layer_order.index(n))
            group = []
            for n in names:
                sl = ShardLayer(self.layer_order.index(n), shard_name, n, (shard_name, n) in self.written_shard_layers)
                sl.layer_number                                # validates the name like the reference does
                group.append(sl)
            yield group

    @classmethod
    def like_model(cls, model_path: Path, output_path: Path, output_astype: torch.dtype = torch.bfloat16):
        index_path = Path(model_path) / "model.safetensors.index.json"
        if not index_path.exists():
            raise FileNotFoundError(f"Model index not found at {index_path}")
        with open(index_path) as fh:
            base_index = json.load(fh)
        order = []
        for file in Path(model_path).glob("*.safetensors"):
            with safe_open(file, framework="pt") as f:
                order.extend(f.keys())
        return cls(base_index=base_index, output_path=Path(output_path), layer_order=order, output_astype=output_astype)
