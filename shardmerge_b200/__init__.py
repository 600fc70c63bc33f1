"""shardmerge_b200 -- B200-native (sm_100a) implementation of ShardMerge's per-tensor spectral
merge hot path, behind the reference's own Python signatures.

    shardmerge_b200.tensor.functions   <->  shard/tensor/functions.py
    shardmerge_b200.merge.fast_fourier <->  shard/merge/fast_fourier.py (FourierMerge)
    shardmerge_b200.merge.base         <->  shard/merge/base.py (MergeTensorsBase)
    shardmerge_b200.writer             <->  shard/writer.py (ModelWriter, ShardLayer)
    shardmerge_b200.config/constants   <->  shard/config.py, shard/constants.py (schema only)

The math runs in hand-written CUDA kernels (shardmerge_b200/csrc) reached through the C ABI
of include/shardmerge_b200.h; there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401  (load lazily on first use; import must work without a GPU)
