"""make_golden_cli.py -- files-in / files-out fixture: run the REFERENCE CLI (`python -m shard merge cfg.yaml`,
shard/__main__.py:47-158, read-only from /root/reference, device cpu) on a small synthetic model laid out on disk the
way the reference's DownloadManager leaves models (shard/index.py:88-95), and keep its output shards under
tests/golden/cli_tiny/.  tests/test_gpu_pipeline.py regenerates the same input files from the same seeds
(`write_models`), runs shardmerge_b200's FourierMerge.merge("cuda") on them and compares tensor by tensor.

    python oracle/make_golden_cli.py          # here: the reference does not travel to the GPU box

TEST INFRASTRUCTURE, not product code.  The model keeps TinyLlama's tensor names and roles (BASELINE config 1) with
reduced row counts so the fixture stays small; the hidden size stays 2048 because the reference degenerates (NaN
imaginary path -> merged delta = 0) on 1-D tensors shorter than that (caveat).
"""
from __future__ import annotations

import json
import os
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

import torch
from safetensors.torch import save_file

REF = "/root/reference"
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "cli_tiny"
H, Q, KV, I, V, L = 2048, 48, 16, 44, 8, 2
ALPHAS = (0.3, 0.5)
SIGMAS = (0.002, 0.0026)


def tensor_specs():
    specs = [("model.embed_tokens.weight", (V, H))]
    for l in range(L):
        p = f"model.layers.{l}."
        specs += [(p + "input_layernorm.weight", (H,)), (p + "mlp.down_proj.weight", (H, I)), (p + "mlp.gate_proj.weight", (I, H)),
                  (p + "mlp.up_proj.weight", (I, H)), (p + "post_attention_layernorm.weight", (H,)),
                  (p + "self_attn.k_proj.weight", (KV, H)), (p + "self_attn.o_proj.weight", (H, Q)),
                  (p + "self_attn.q_proj.weight", (Q, H)), (p + "self_attn.v_proj.weight", (KV, H))]
    specs += [("model.norm.weight", (H,)), ("lm_head.weight", (V, H))]
    return specs


def make_models():
    """-> {model name: {tensor name: bf16 tensor}} for synth/base, synth/ft0, synth/ft1 (CPU generator: same bits anywhere)."""
    models = {"synth/base": {}, "synth/ft0": {}, "synth/ft1": {}}
    for idx, (name, shape) in enumerate(tensor_specs()):
        g = torch.Generator().manual_seed(1234 + idx)
        if len(shape) == 1:
            base = (1.0 + 0.1 * torch.randn(shape, generator=g)).to(torch.bfloat16)
            sig = (0.01, 0.013)
        else:
            base = (0.02 * torch.randn(shape, generator=g)).to(torch.bfloat16)
            sig = SIGMAS
        models["synth/base"][name] = base
        for k in range(2):
            gk = torch.Generator().manual_seed(100000 * (k + 1) + idx)
            models[f"synth/ft{k}"][name] = (base.float() + sig[k] * torch.randn(shape, generator=gk)).to(torch.bfloat16)
    return models


def shard_of(name: str) -> str:
    if name.startswith("model.layers.1.") or name in ("model.norm.weight", "lm_head.weight"):
        return "model-00002-of-00002.safetensors"
    return "model-00001-of-00002.safetensors"


def write_models(storage: Path):
    for model, tensors in make_models().items():
        d = storage / model
        d.mkdir(parents=True, exist_ok=True)
        wm = {n: shard_of(n) for n in tensors}
        for shard in sorted(set(wm.values())):
            save_file({n: t for n, t in tensors.items() if wm[n] == shard}, str(d / shard), metadata={"format": "pt"})
        (d / "model.safetensors.index.json").write_text(json.dumps(
            {"metadata": {"total_size": sum(t.numel() * 2 for t in tensors.values())}, "weight_map": wm}, indent=2))


def config_yaml(storage: Path, cache: Path, out: Path, device: str) -> str:
    return (f'output_base_model: "synth/base"\noutput_dtype: "bfloat16"\nfinetune_merge:\n'
            f'  - {{ "model": "synth/ft0", "base": "synth/base", "alpha": {ALPHAS[0]}, "is_input": true }}\n'
            f'  - {{ "model": "synth/ft1", "base": "synth/base", "alpha": {ALPHAS[1]}, "is_output": true }}\n'
            f'output_dir: "{out}"\ndevice: "{device}"\nclean_cache: false\ncache_dir: "{cache}"\nstorage_dir: "{storage}"\n')


def main():
    with tempfile.TemporaryDirectory() as td:
        td = Path(td)
        write_models(td / "storage")
        (td / "cfg.yaml").write_text(config_yaml(td / "storage", td / "cache", td / "out", "cpu"))
        env = dict(os.environ, PYTHONPATH=REF)
        subprocess.run([sys.executable, "-m", "shard", "merge", str(td / "cfg.yaml")], cwd=td, env=env, check=True)
        if OUT.exists():
            shutil.rmtree(OUT)
        OUT.mkdir(parents=True)
        for f in sorted((td / "out").iterdir()):
            shutil.copy(f, OUT / f.name)
        (OUT / "MANIFEST.txt").write_text(
            f"output of `python -m shard merge` ({REF}, torch {torch.__version__}, device cpu) on the files oracle/make_golden_cli.py:"
            f"write_models() generates; H={H} Q={Q} KV={KV} I={I} V={V} L={L}\n")
    # sanity: how much of every merged tensor differs from the base (a degenerate reference would return the base)
    from safetensors import safe_open
    base = make_models()["synth/base"]
    for f in sorted(OUT.glob("*.safetensors")):
        with safe_open(f, framework="pt") as sf:
            for k in sf.keys():
                t = sf.get_tensor(k)
                print(f"{f.name} {k} {tuple(t.shape)} differs-from-base {float((t != base[k]).float().mean()):.3f}")


if __name__ == "__main__":
    main()
