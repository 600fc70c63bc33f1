"""Local tensor sources with the part of HFMultiModelIndex's interface (shard/index.py:60-276)
that the merge path calls: add_model, model_indexes, get_model_keys, get_layer_order,
get_tensor -> promise, preload_tensor.  Network download, claim counting and the HF cache
layout are out of scope; a reference HFMultiModelIndex instance
can be passed to FourierMerge instead of these classes.

  InMemoryIndex        tensors handed over as dicts (tests, benchmarks, synthetic models)
  LocalSafetensorsIndex {storage_dir}/{org}/{model}/model.safetensors.index.json + shards on disk,
                       read lazily, uploaded with pinned staging, nothing kept after use
                       (the reference keeps every tensor it ever read in RAM, index.py:79,265).
"""
from __future__ import annotations

import json
import re
from pathlib import Path
from typing import Dict, List, Set

import torch


class TensorPromise:
    """Awaitable handle for one tensor (interface of shard/index.py:38-58)."""

    def __init__(self, model_uri: str, tensor_name: str, device: str, loader):
        self.model_uri, self.tensor_name, self.device = model_uri, tensor_name, device
        self._loader = loader

    async def get(self) -> torch.Tensor:
        return self._loader()


def canonical_layer_order(names) -> List[str]:
    """embed_tokens, then model.layers.<n>.<component> by (n, component name as discovered on layer 0),
    then model.norm.weight, lm_head, then everything else sorted (shard/index.py:132-187)."""
    names = list(names)
    embed = sorted(n for n in names if "embed_tokens" in n)
    layer = [n for n in names if "layers." in n]
    norm = sorted(n for n in names if "model.norm.weight" in n)
    head = sorted(n for n in names if "lm_head" in n)
    claimed = set(embed) | set(layer) | set(norm) | set(head)
    other = sorted(n for n in names if n not in claimed)
    numbers = sorted({int(n.split("layers.")[1].split(".")[0]) for n in layer})
    prefix0 = "model.layers.0."
    components = sorted(n[len(prefix0):] for n in layer if n.startswith(prefix0))
    ordered_layers = [f"model.layers.{i}.{c}" for i in numbers for c in components]
    ordered = embed + ordered_layers + norm + head + other
    if set(ordered) != set(names):
        raise ValueError(f"Weight ordering mismatch! Missing: {set(names) - set(ordered)}, Extra: {set(ordered) - set(names)}")
    return ordered


class _IndexBase:
    def __init__(self):
        self.model_indexes: Dict[str, dict] = {}
        self._order: Dict[str, List[str]] = {}
        self._prefetched: dict = {}
        self._copy_streams: dict = {}

    def _register(self, model_uri: str, index: dict):
        self.model_indexes[model_uri] = index
        self._order[model_uri] = canonical_layer_order(index["weight_map"].keys())

    def get_model_keys(self, model_uri: str) -> Set[str]:
        return set(self.model_indexes[model_uri]["weight_map"].keys())

    def get_layer_order(self, model_uri: str) -> List[str]:
        return list(self._order[model_uri])

    async def preload_tensor(self, model_uri: str, tensor_name: str):
        return None

    # ---- one-tensor-ahead upload ------------------------------------------------------
    # The reference loads a tensor when _merge_layer asks for it, synchronously (shard/index.py:238-270), so the
    # GPU idles during every host-to-device copy.  prefetch() starts the copy of a tensor the merge loop will ask
    # for next on a dedicated copy stream; the matching get_tensor(...).get() hands the device tensor over after
    # making the caller's stream wait for the copy.
    def _host_tensor(self, model_uri: str, tensor_name: str) -> torch.Tensor:
        raise NotImplementedError

    def prefetch(self, model_uri: str, tensor_name: str, device: str):
        dev = torch.device(device)
        if dev.type != "cuda":
            return
        key = (model_uri, tensor_name, str(dev))
        if key in self._prefetched:
            return
        host = self._host_tensor(model_uri, tensor_name)
        if host.device.type == "cuda":
            return
        if not host.is_pinned():
            host = host.pin_memory()
        stream = self._copy_streams.get(str(dev))
        if stream is None:
            stream = self._copy_streams[str(dev)] = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(stream):
            t = host.to(dev, non_blocking=True)
            ev = stream.record_event()
        self._prefetched[key] = (t, ev, host)

    def _take_prefetched(self, model_uri: str, tensor_name: str, device: str):
        ent = self._prefetched.pop((model_uri, tensor_name, str(torch.device(device))), None)
        if ent is None:
            return None
        t, ev, _host = ent
        cur = torch.cuda.current_stream(t.device)
        cur.wait_event(ev)
        t.record_stream(cur)
        return t


class InMemoryIndex(_IndexBase):
    """models: {model_uri: {tensor_name: tensor}}; `shard_of(tensor_name) -> file name` lays out the
    output shards (default: one shard per transformer layer, like HF checkpoints roughly do)."""

    def __init__(self, models: Dict[str, Dict[str, torch.Tensor]], shard_of=None):
        super().__init__()
        self.models = models
        self._shard_of = shard_of or _default_shard_of

    async def add_model(self, model_uri: str, revision: str = "main"):
        if model_uri in self.model_indexes:
            return
        tensors = self.models[model_uri]
        self._register(model_uri, {"metadata": {"total_size": sum(t.numel() * t.element_size() for t in tensors.values())},
                                   "weight_map": {n: self._shard_of(n) for n in tensors}})

    def tensor_numels(self, model_uri: str) -> Dict[str, int]:
        return {n: t.numel() for n, t in self.models[model_uri].items()}

    def tensor_shapes(self, model_uri: str) -> Dict[str, tuple]:
        return {n: tuple(t.shape) for n, t in self.models[model_uri].items()}

    def _host_tensor(self, model_uri: str, tensor_name: str) -> torch.Tensor:
        return self.models[model_uri][tensor_name]

    def get_tensor(self, model_uri: str, tensor_name: str, device: str = "cpu") -> TensorPromise:
        t = self.models[model_uri][tensor_name]

        def load():
            if str(t.device) == str(device):
                return t
            pre = self._take_prefetched(model_uri, tensor_name, device)
            return pre if pre is not None else t.to(device, non_blocking=t.is_pinned())

        return TensorPromise(model_uri, tensor_name, device, load)


def _default_shard_of(name: str) -> str:
    m = re.match(r"model\.layers\.(\d+)\.", name)
    return f"model-layer{int(m.group(1)):05d}.safetensors" if m else "model-misc.safetensors"


class LocalSafetensorsIndex(_IndexBase):
    """Models stored under storage_dir/<model_uri>/ as model.safetensors.index.json + shard files
    (the layout the reference's DownloadManager leaves behind, shard/index.py:88-95)."""

    def __init__(self, storage_dir):
        super().__init__()
        self.storage_dir = Path(storage_dir)

    async def add_model(self, model_uri: str, revision: str = "main"):
        if model_uri in self.model_indexes:
            return
        path = self.storage_dir / model_uri / "model.safetensors.index.json"
        if not path.exists():
            raise FileNotFoundError(f"{path} not found (shardmerge_b200 does not download models)")
        with open(path) as fh:
            self._register(model_uri, json.load(fh))

    def tensor_shapes(self, model_uri: str) -> Dict[str, tuple]:
        """shapes from the safetensors headers (no tensor data is read)."""
        from safetensors import safe_open
        out: Dict[str, tuple] = {}
        index = self.model_indexes[model_uri]
        for shard in sorted(set(index["weight_map"].values())):
            with safe_open(str(self.storage_dir / model_uri / shard), framework="pt") as f:
                for key in f.keys():
                    out[key] = tuple(f.get_slice(key).get_shape())
        return out

    def tensor_numels(self, model_uri: str) -> Dict[str, int]:
        """element counts from the safetensors headers (no tensor data is read)."""
        out: Dict[str, int] = {}
        for key, shape in self.tensor_shapes(model_uri).items():
            n = 1
            for d in shape:
                n *= d
            out[key] = n
        return out

    def get_tensor(self, model_uri: str, tensor_name: str, device: str = "cpu") -> TensorPromise:
        index = self.model_indexes[model_uri]
        if tensor_name not in index["weight_map"]:
            raise KeyError(f"Tensor {tensor_name} not found in model {model_uri}")
        shard = self.storage_dir / model_uri / index["weight_map"][tensor_name]

        def load():
            if torch.device(device).type == "cuda":
                pre = self._take_prefetched(model_uri, tensor_name, device)
                if pre is not None:
                    return pre
                return self._host_tensor(model_uri, tensor_name).pin_memory().to(device, non_blocking=True)
            return self._host_tensor(model_uri, tensor_name)

        return TensorPromise(model_uri, tensor_name, device, load)

    def _host_tensor(self, model_uri: str, tensor_name: str) -> torch.Tensor:
        from safetensors import safe_open
        index = self.model_indexes[model_uri]
        shard = self.storage_dir / model_uri / index["weight_map"][tensor_name]
        with safe_open(str(shard), framework="pt") as f:
            return f.get_tensor(tensor_name)
