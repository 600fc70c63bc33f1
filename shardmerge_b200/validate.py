"""Output validation and model-file copying.

  verify_output     what scripts/verify_safetensors.py does for a merged directory -- every shard named by
                    model.safetensors.index.json exists, holds exactly the tensors the index assigns to it, nothing extra
                    -- plus what a merge can get wrong beyond the index: dtype, shape and non-finite values per tensor.
  copy_model_files  the local counterpart of `python -m shard copy-model` (shard/__main__.py:160-201, which pulls a
                    model's configuration / tokenizer files from the Hugging Face hub into the output directory): the
                    same files, copied from a model directory on disk (no network here).
Host-side utilities: they read files, they do not touch the merge path.
"""
from __future__ import annotations

import json
import shutil
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, List, Optional

import torch
from safetensors import safe_open

_WEIGHT_SUFFIXES = (".safetensors", ".bin", ".pt", ".pth", ".gguf", ".h5", ".msgpack")


@dataclass
class VerifyReport:
    missing_files: List[str] = field(default_factory=list)
    extra_files: List[str] = field(default_factory=list)
    missing_keys: Dict[str, List[str]] = field(default_factory=dict)      # in the index but not in the file
    extra_keys: Dict[str, List[str]] = field(default_factory=dict)        # in the file but not in the index
    wrong_dtype: Dict[str, str] = field(default_factory=dict)
    wrong_shape: Dict[str, tuple] = field(default_factory=dict)
    non_finite: Dict[str, int] = field(default_factory=dict)
    tensors: int = 0

    @property
    def ok(self) -> bool:
        return not (self.missing_files or self.extra_files or self.missing_keys or self.extra_keys or self.wrong_dtype
                    or self.wrong_shape or self.non_finite)

    def summary(self) -> str:
        if self.ok:
            return f"All safetensors files align with the index ({self.tensors} tensors, no non-finite values)"
        parts = []
        for label, val in (("missing files", self.missing_files), ("extra files", self.extra_files),
                           ("files with missing keys", self.missing_keys), ("files with extra keys", self.extra_keys),
                           ("tensors with a wrong dtype", self.wrong_dtype), ("tensors with a wrong shape", self.wrong_shape),
                           ("tensors with NaN / Inf", self.non_finite)):
            if val:
                parts.append(f"{len(val)} {label}: {sorted(val)[:5]}")
        return "; ".join(parts)


def verify_output(model_dir, expected_dtype: Optional[torch.dtype] = None, expected_shapes: Optional[Dict[str, tuple]] = None,
                  check_values: bool = True, device: str = "cpu") -> VerifyReport:
    """Check a merged model directory against its own model.safetensors.index.json."""
    model_dir = Path(model_dir)
    with open(model_dir / "model.safetensors.index.json") as fh:
        weight_map = json.load(fh)["weight_map"]
    expected: Dict[str, set] = {}
    for key, file in weight_map.items():
        expected.setdefault(file, set()).add(key)
    present = {p.name for p in model_dir.glob("*.safetensors")}
    rep = VerifyReport(missing_files=sorted(set(expected) - present), extra_files=sorted(present - set(expected)))
    for file in sorted(set(expected) & present):
        with safe_open(str(model_dir / file), framework="pt") as f:
            keys = set(f.keys())
            if expected[file] - keys:
                rep.missing_keys[file] = sorted(expected[file] - keys)
            if keys - expected[file]:
                rep.extra_keys[file] = sorted(keys - expected[file])
            for key in sorted(keys & expected[file]):
                rep.tensors += 1
                sl = f.get_slice(key)
                shape = tuple(sl.get_shape())
                if expected_shapes is not None and key in expected_shapes and tuple(expected_shapes[key]) != shape:
                    rep.wrong_shape[key] = shape
                if expected_dtype is None and not check_values:
                    continue
                t = f.get_tensor(key)
                if expected_dtype is not None and t.dtype != expected_dtype:
                    rep.wrong_dtype[key] = str(t.dtype)
                if check_values and t.is_floating_point():
                    bad = int((~torch.isfinite(t.to(device))).sum().item())
                    if bad:
                        rep.non_finite[key] = bad
    return rep


def copy_model_files(model_dir, output_dir) -> List[str]:
    """Copy everything that is not a weight file or a weight index (config.json, generation_config.json, tokenizer files,
    ...) from a model directory into the output directory; existing files are left alone.  -> names copied."""
    model_dir, output_dir = Path(model_dir), Path(output_dir)
    if not model_dir.is_dir():
        raise FileNotFoundError(f"model directory {model_dir} not found (shardmerge_b200 does not download models)")
    output_dir.mkdir(parents=True, exist_ok=True)
    copied = []
    for src in sorted(model_dir.iterdir()):
        if not src.is_file() or src.name.endswith(_WEIGHT_SUFFIXES) or src.name.endswith(".index.json"):
            continue
        dst = output_dir / src.name
        if not dst.exists():
            shutil.copy2(src, dst)
            copied.append(src.name)
    return copied
