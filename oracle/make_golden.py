"""make_golden.py -- generate tests/golden/*.npz by running the REFERENCE ITSELF (read-only,
imported from /root/reference) on seeded inputs, CPU, torch as installed in this container.

Run here (the reference does not travel to the GPU box):
    python oracle/make_golden.py
The fixtures pin oracle/oracle_np.py (tests/test_oracle_golden.py) and are compared against
the CUDA path on the GPU (tests/test_gpu_parity.py).  TEST INFRASTRUCTURE, not product code.
"""
from __future__ import annotations

import asyncio
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

REF = "/root/reference"
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def bf16_bits(t: torch.Tensor) -> np.ndarray:
    return t.contiguous().view(torch.int16).numpy().view(np.uint16).copy()


def synth(shape, seed, n_models, sigmas=(0.002, 0.0026, 0.0023, 0.0029), base_sigma=0.02, base_mean=0.0):
    """Synthetic weights as in SURVEY.md section 8d (CPU generator here: fixtures carry the bits)."""
    g = torch.Generator().manual_seed(1234 + seed)
    base = (base_mean + base_sigma * torch.randn(shape, generator=g)).to(torch.bfloat16)
    fts = []
    for k in range(n_models):
        gk = torch.Generator().manual_seed(100000 * (k + 1) + seed)
        fts.append((base.float() + sigmas[k] * torch.randn(shape, generator=gk)).to(torch.bfloat16))
    return base, fts


def tensor_level_cases(F):
    """merge_tensors_fft2_slerp with stage taps (shard/tensor/functions.py:164-221)."""
    cases = {}
    specs = [("v2048", (2048,), 11), ("m32x64", (32, 64), 12), ("m64x128", (64, 128), 13),
             ("m96x40", (96, 40), 14), ("m256x128", (256, 128), 15)]
    for name, shape, seed in specs:
        g = torch.Generator().manual_seed(seed)
        v0 = 0.0026 * torch.randn(shape, generator=g)
        v1 = 0.0020 * torch.randn(shape, generator=g)
        taps = {}
        orig = F.interpolate_fft_components

        def spy(v0_fft, v1_fft, *a, **kw):
            outer = "fft0" not in taps
            if outer:
                taps["fft0"] = v0_fft.clone(); taps["fft1"] = v1_fft.clone()
            r = orig(v0_fft, v1_fft, *a, **kw)
            if outer:
                taps["res"] = r.clone()
            return r

        F.interpolate_fft_components = spy
        try:
            merged, n0, n1 = F.merge_tensors_fft2_slerp(v0, v1, t=0.375, device="cpu", t_sum=1.0, cutoff_pct=0.08,
                                                        cull_pct=0.20)
        finally:
            F.interpolate_fft_components = orig
        # the same blend with interp_imag=False (functions.py:160) isolates the real-part blend
        res_noimag = F.interpolate_fft_components(taps["fft0"], taps["fft1"], t=0.375, device="cpu", t_sum=1.0,
                                                  cutoff_pct=0.08, cull_pct=0.20, interp_imag=False)
        cases[name] = dict(v0=v0.numpy(), v1=v1.numpy(), t=np.float64(0.375), merged=merged.numpy(),
                           n0=np.float64(n0), n1=np.float64(n1),
                           fft0=taps["fft0"].numpy(), fft1=taps["fft1"].numpy(), res=taps["res"].numpy(),
                           res_real_noimag=res_noimag.real.numpy().copy())
    return cases


class _Promise:
    def __init__(self, t):
        self.t = t

    async def get(self):
        return self.t


class _StubIndex:
    """Stands in for HFMultiModelIndex (shard/index.py:195-236) with in-memory tensors."""

    def __init__(self, tensors):
        self.tensors = tensors

    def get_tensor(self, model, layer_name, device="cpu"):
        return _Promise(self.tensors[(model, layer_name)].to(device))

    async def preload_tensor(self, model, layer_name):
        return None


def layer_level_cases():
    """FourierMerge._merge_layer (shard/merge/fast_fourier.py:103-276) on synthetic bf16 models."""
    from shard.config import MergeConfig, MergeModel
    from shard.merge.fast_fourier import FourierMerge
    from shard.writer import ShardLayer

    cases = {}

    def run(name, shape, seed, alphas, sigmas=None, n_models=None, mutate=None, layer="model.layers.3.mlp.up_proj.weight",
            flags=None, base_mean=0.0, base_sigma=0.02):
        n_models = n_models or len(alphas)
        base, fts = synth(shape, seed, n_models, sigmas or (0.002, 0.0026, 0.0023, 0.0029), base_sigma, base_mean)
        if mutate:
            base, fts = mutate(base, fts)
        tensors = {("org/base", layer): base}
        models = []
        for k, ft in enumerate(fts):
            tensors[(f"org/ft{k}", layer)] = ft
            kw = dict(model=f"org/ft{k}", base="org/base", alpha=alphas[k])
            if flags and k in flags:
                kw.update(flags[k])
            models.append(MergeModel(**kw))
        with tempfile.TemporaryDirectory() as td:
            cfg = MergeConfig(finetune_merge=models, output_base_model="org/base", output_dir=td + "/out",
                              cache_dir=td + "/cache", storage_dir=td + "/st")
            merger = FourierMerge(cfg, index_manager=_StubIndex(tensors))
            out = asyncio.run(merger._merge_layer(ShardLayer(0, "s", layer, False), "cpu"))
        d = dict(base=bf16_bits(base), alphas=np.array(alphas, dtype=np.float64), layer=np.array(layer))
        for k, ft in enumerate(fts):
            d[f"ft{k}"] = bf16_bits(ft)
        if out.dtype == torch.bfloat16:
            d["out"] = bf16_bits(out)
        else:
            d["out_f32"] = out.float().numpy()
        cases[name] = d

    run("slerp_256x512", (256, 512), 1, [0.3, 0.5])
    run("slerp_swapped_128x256", (128, 256), 2, [0.3, 0.5], sigmas=(0.0026, 0.002))
    run("slerp_a7b2_128x256", (128, 256), 3, [0.7, 0.2])
    run("slerp_1d_2048", (2048,), 4, [0.3, 0.5], sigmas=(0.01, 0.013), base_mean=1.0, base_sigma=0.1,
        layer="model.layers.3.input_layernorm.weight")
    run("slerp_352x96", (352, 96), 5, [0.3, 0.5])                 # 352 = 2^5*11, 48 = 2^4*3
    run("tree4_128x256", (128, 256), 6, [0.3, 0.5, 0.4, 0.2])
    run("tree3_64x256", (64, 256), 7, [0.3, 0.5, 0.4])
    run("arith_128x256", (128, 256), 8, [0.3, 0.5], sigmas=(0.0026, 0.0001))
    run("single_64x128", (64, 128), 9, [0.3])
    run("add_zero_64x128", (64, 128), 10, [0.3, 0.5], mutate=lambda b, f: (b, [b.clone(), b.clone()]))
    run("onezero_64x128", (64, 128), 16, [0.3, 0.5], mutate=lambda b, f: (b, [f[0], b.clone()]))
    run("layer_range_64x128", (64, 128), 17, [0.3, 0.5, 0.4], flags={2: dict(start_layer=10)})
    run("slerp_oddC_64x129", (64, 129), 18, [0.3, 0.5])           # odd row length (the CUDA path merges the transpose)
    return cases


def main():
    sys.path.insert(0, REF)
    import shard.tensor.functions as F

    torch.set_num_threads(4)
    OUT.mkdir(parents=True, exist_ok=True)
    tl = tensor_level_cases(F)
    for name, d in tl.items():
        np.savez_compressed(OUT / f"tensor_{name}.npz", **d)
    ll = layer_level_cases()
    for name, d in ll.items():
        np.savez_compressed(OUT / f"layer_{name}.npz", **d)
    with open(OUT / "MANIFEST.txt", "w") as f:
        f.write(f"generated by oracle/make_golden.py from {REF} with torch {torch.__version__} on CPU\n")
        for name in sorted(list(tl) + list(ll)):
            f.write(name + "\n")
    print("wrote", len(tl) + len(ll), "fixtures to", OUT)


if __name__ == "__main__":
    main()
