"""Files in -> files out through the public merge() API (FourierMerge + LocalSafetensorsIndex + ModelWriter), wall clock:
    python tools/bench_files.py [--layers 4] [--workload llama8b] [--root /dev/shm]
Writes a synthetic base + 2 finetunes (L layers of the named architecture, one safetensors shard per layer) under ROOT,
merges them into ROOT/out with `device: cuda`, verifies the output (validate.verify_output), prints merged params/s.
Everything a user's `python -m shard merge` run does except the download: file reads, uploads, kernels, downloads, file writes."""
import argparse, asyncio, json, shutil, sys, tempfile, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench as B
from safetensors.torch import save_file
from shardmerge_b200.config import MergeConfig, MergeModel
from shardmerge_b200.index import LocalSafetensorsIndex
from shardmerge_b200.merge.fast_fourier import FourierMerge
from shardmerge_b200.validate import verify_output

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=4)
ap.add_argument("--workload", default="llama8b")
ap.add_argument("--root", default="/dev/shm")
ap.add_argument("--repeats", type=int, default=2)
ap.add_argument("--reader-threads", type=int, default=4)
args = ap.parse_args()
a = B.ARCH[args.workload]
dev = torch.device("cuda:0")
root = Path(tempfile.mkdtemp(prefix="shardmerge_files_", dir=args.root))
try:
    per_layer = B.layer_tensors(a)
    models = ["synth/base", "synth/ft0", "synth/ft1"]
    wm, idx, params = {}, 0, 0
    for m in models:
        (root / "storage" / m).mkdir(parents=True)
    for layer in range(args.layers):
        shard = f"model-{layer + 1:05d}-of-{args.layers:05d}.safetensors"
        blobs = {m: {} for m in models}
        for nm, shape in per_layer:
            name = f"model.layers.{layer}.{nm}"
            base, fts = B.synth_tensor(torch, shape, idx, 2, dev)
            blobs["synth/base"][name] = base.cpu(); blobs["synth/ft0"][name] = fts[0].cpu(); blobs["synth/ft1"][name] = fts[1].cpu()
            wm[name] = shard; idx += 1; params += B.numel(shape)
        for m in models:
            save_file(blobs[m], str(root / "storage" / m / shard), metadata={"format": "pt"})
    for m in models:
        (root / "storage" / m / "model.safetensors.index.json").write_text(json.dumps({"metadata": {}, "weight_map": wm}))
    in_bytes = 3 * params * 2
    rates = []
    for rep in range(args.repeats):
        out = root / f"out{rep}"
        cfg = MergeConfig(finetune_merge=[MergeModel(model="synth/ft0", base="synth/base", alpha=0.3, is_input=True),
                                          MergeModel(model="synth/ft1", base="synth/base", alpha=0.5, is_output=True)],
                          output_base_model="synth/base", output_dir=str(out), device=str(dev), storage_dir=str(root / "storage"))
        fm = FourierMerge(cfg, index_manager=LocalSafetensorsIndex(cfg.storage_path, reader_threads=args.reader_threads))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        asyncio.run(fm.merge(str(dev)))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        rates.append(params / dt)
        rep_ok = verify_output(out, expected_dtype=torch.bfloat16).ok
        print(f"run {rep}: {dt:.3f} s  {params / dt / 1e9:.2f} Gparam/s  (read {in_bytes / dt / 1e9:.1f} GB/s, written {params * 2 / dt / 1e9:.1f} GB/s)  verified={rep_ok}")
        shutil.rmtree(out)
    print(json.dumps(dict(workload=f"{args.workload}-shaped, {args.layers} layers, 2 finetunes, safetensors on {args.root} in and out",
                          merged_params=params, params_per_s=max(rates), api="FourierMerge.merge() + LocalSafetensorsIndex + ModelWriter")))
finally:
    shutil.rmtree(root, ignore_errors=True)
