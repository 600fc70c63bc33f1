"""ModelWriter / ShardLayer with the interface of shard/writer.py.

Observable contract kept from the reference: the output directory holds a copy of the base
model's model.safetensors.index.json (:75-81), one safetensors file per entry of its weight_map
with metadata {"format": "pt"} (:143), tensors cast to output_astype (:133); existing shards are
scanned at start and their tensors count as written (resume, :93-113); finalize() raises if a
tensor is missing (:151-161).

What changes is the I/O schedule.  The reference re-reads and re-writes the
whole shard for every tensor (:125-143, O(k^2) bytes per shard) and copies device -> host
synchronously (:133).  Here

  * a tensor is copied device -> pinned host memory asynchronously on a side stream
    (`_stage`); the pinned buffers come from a pool (no cudaHostAlloc per tensor) and go back to
    it when their shard is on disk;
  * a shard is written exactly once, when its last tensor has arrived, by a worker thread
    (`_ShardWriter`) so the file write overlaps the merge of the next shard.  The file is the
    safetensors container (8-byte little-endian header length, JSON header, raw little-endian
    tensor data) written straight from the pinned buffers, byte-identical to what
    `safetensors.torch.save_file` produces for the same tensors, to a temporary name that is
    renamed into place (a reader never sees a partial shard);
  * `flush_partial()` writes the incomplete shards too (called when a merge aborts), so a
    resumed run continues at tensor granularity like the reference's.
"""
from __future__ import annotations

import json
import logging
import mmap
import os
import queue
import struct
import threading
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, Generator, List, Optional, Set

import torch
from safetensors import safe_open

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


# ------------------------------------------------------------------------------------------
# safetensors container, written from host buffers without an intermediate copy
# ------------------------------------------------------------------------------------------
# dtype -> (safetensors name, rank in the format's dtype order: tensors are laid out by descending rank, then name)
_ST_DTYPES = {
    torch.bool: ("BOOL", 0), torch.uint8: ("U8", 1), torch.int8: ("I8", 2),
    torch.float8_e5m2: ("F8_E5M2", 3), torch.float8_e4m3fn: ("F8_E4M3", 4),
    torch.int16: ("I16", 5), torch.uint16: ("U16", 6), torch.float16: ("F16", 7), torch.bfloat16: ("BF16", 8),
    torch.int32: ("I32", 9), torch.uint32: ("U32", 10), torch.float32: ("F32", 11), torch.float64: ("F64", 12),
    torch.int64: ("I64", 13), torch.uint64: ("U64", 14),
}


_CHUNK = 32 << 20          # bytes per positional write handed to the I/O pool


def write_safetensors(path, tensors: Dict[str, torch.Tensor], metadata: Optional[Dict[str, str]] = None, io_pool=None):
    """Write `tensors` (contiguous CPU tensors) as one safetensors file.  Same bytes as
    safetensors.torch.save_file(tensors, path, metadata) (tests/test_host_logic.py checks that), but the data goes
    from the tensors' own (pinned) memory to the file: no per-tensor bytes() copy.  With `io_pool` (a
    concurrent.futures executor) the data is written as positional 32 MB chunks in parallel -- a single writer thread
    moves ~3 GB/s into the page cache, a merge on one B200 produces ~15 GB/s."""
    items = []
    for name, t in tensors.items():
        if t.device.type != "cpu" or not t.is_contiguous():
            raise ValueError(f"write_safetensors needs contiguous CPU tensors ({name})")
        if t.dtype not in _ST_DTYPES:
            raise ValueError(f"dtype {t.dtype} of {name} has no safetensors name")
        items.append((name, t))
    items.sort(key=lambda it: (-_ST_DTYPES[it[1].dtype][1], it[0]))
    header = {}
    if metadata is not None:
        header["__metadata__"] = dict(metadata)
    off = 0
    for name, t in items:
        n = t.numel() * t.element_size()
        header[name] = {"dtype": _ST_DTYPES[t.dtype][0], "shape": list(t.shape), "data_offsets": [off, off + n]}
        off += n
    blob = json.dumps(header, separators=(",", ":"), ensure_ascii=False).encode("utf-8")
    blob += b" " * (-len(blob) % 8)
    head = struct.pack("<Q", len(blob)) + blob
    total = len(head) + off
    fd = os.open(str(path), os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o644)
    try:
        if io_pool is None or total < (4 << 20):
            os.pwrite(fd, head, 0)
            pos = len(head)
            for _, t in items:
                n = t.numel() * t.element_size()
                if n:
                    _pwrite_all(fd, memoryview(t.reshape(-1).view(torch.uint8).numpy()), pos)
                pos += n
        else:
            # writes to ONE file serialise on its inode lock; stores through a shared mapping do not: map the file and
            # let the pool copy 32 MB chunks in parallel (torch's copy releases the GIL)
            os.ftruncate(fd, total)
            mm = mmap.mmap(fd, total)
            try:
                mm[: len(head)] = head
                dst = torch.frombuffer(mm, dtype=torch.uint8)
                pos = len(head)
                futures = []
                for _, t in items:
                    n = t.numel() * t.element_size()
                    if n:
                        src = t.reshape(-1).view(torch.uint8)
                        for c0 in range(0, n, _CHUNK):
                            c1 = min(c0 + _CHUNK, n)
                            futures.append(io_pool.submit(dst[pos + c0: pos + c1].copy_, src[c0:c1]))
                    pos += n
                for f in futures:
                    f.result()
                del dst, futures
            finally:
                mm.close()
    finally:
        os.close(fd)


def _pwrite_all(fd: int, data, offset: int):
    done = 0
    while done < len(data):
        done += os.pwrite(fd, data[done:], offset + done)


class _PinnedPool:
    """Pinned host buffers by byte size; a merge touches a handful of distinct tensor sizes, so after the first
    shard every `_stage` re-uses a buffer instead of calling cudaHostAlloc."""

    def __init__(self):
        self.free: Dict[int, list] = {}
        self.lock = threading.Lock()

    def take(self, nbytes: int) -> torch.Tensor:
        with self.lock:
            lst = self.free.get(nbytes)
            if lst:
                return lst.pop()
        return torch.empty(max(nbytes, 1), dtype=torch.uint8).pin_memory()

    def give(self, buf: torch.Tensor):
        with self.lock:
            self.free.setdefault(buf.numel(), []).append(buf)


class _ShardWriter:
    """Worker threads that turn (shard, staged tensors) jobs into files: `workers` shards are in progress at a time, their
    data is written as parallel positional chunks by `io_threads` more threads (file writes release the GIL).  At most
    `depth` jobs wait in the queue, so the pinned memory a merge holds is bounded by depth + workers + 1 shards."""

    def __init__(self, depth: int = 4, workers: int = 4, io_threads: int = 8):
        from concurrent.futures import ThreadPoolExecutor
        self.jobs: "queue.Queue" = queue.Queue(maxsize=depth)
        self.errors: list = []
        self.io_pool = ThreadPoolExecutor(max_workers=io_threads, thread_name_prefix="shardmerge-io")
        self.threads = [threading.Thread(target=self._run, name=f"shardmerge-writer-{i}", daemon=True) for i in range(workers)]
        for t in self.threads:
            t.start()

    def _run(self):
        while True:
            job = self.jobs.get()
            try:
                if job is None:
                    return
                job()
            except Exception as exc:           # kept for finalize(): the reference logs and drops the shard
                self.errors.append(exc)
            finally:
                self.jobs.task_done()

    def submit(self, job):
        self.jobs.put(job)

    def drain(self):
        self.jobs.join()

    def close(self):
        for _ in self.threads:
            self.jobs.put(None)
        for t in self.threads:
            t.join()
        self.io_pool.shutdown(wait=True)


@dataclass
class ModelWriter:
    base_index: dict
    output_path: Path
    layer_order: list
    output_astype: torch.dtype
    written_shard_layers: Set[tuple] = field(default_factory=set)
    shard_to_tensors: Dict[str, Set[str]] = field(default_factory=dict)
    only_shards: Optional[Set[str]] = None       # multi-GPU merge: the shards this process owns (schedule.py)
    write_index: bool = True                     # multi-GPU merge: rank 0 writes the index copy, the others only read it
    async_write: bool = True
    pinned_pool: Optional["_PinnedPool"] = None  # share staging buffers between writers (successive merges in one process)
    queue_depth: int = 4                         # shards waiting for a writer thread before add_tensor blocks (bounds pinned memory)
    writer_threads: int = 4                      # shard files written at a time
    io_threads: int = 8                          # threads copying chunks into the files' mappings

    def __post_init__(self):
        self.output_path = Path(self.output_path)
        self.output_path.mkdir(parents=True, exist_ok=True)
        self.index_path = self.output_path / "model.safetensors.index.json"
        if self.index_path.exists():
            logger.info(f"Index already exists: {self.index_path}")
            with open(self.index_path) as fh:
                self.base_index = json.load(fh)
        elif self.write_index:
            tmp = self.index_path.with_name(self.index_path.name + f".tmp{os.getpid()}")
            with open(tmp, "w") as fh:
                json.dump(self.base_index, fh, indent=2)
            os.replace(tmp, self.index_path)     # a concurrent reader sees the old state or the whole file
        self.shard_to_tensors = {}
        for tensor_name, shard_name in self.base_index["weight_map"].items():
            self.shard_to_tensors.setdefault(shard_name, set()).add(tensor_name)
        self._order_pos = {n: i for i, n in enumerate(self.layer_order)}
        self._staged: Dict[str, Dict[str, torch.Tensor]] = {}      # shard -> {tensor: host tensor (view of a pinned buffer)}
        self._bufs: Dict[str, list] = {}                           # shard -> pinned buffers to give back
        self._events: Dict[str, list] = {}
        self._recent_events: list = []                            # device -> host copies since the last wait_staged()
        self._copy_stream = None
        self._pool = self.pinned_pool if self.pinned_pool is not None else _PinnedPool()
        self._worker: Optional[_ShardWriter] = None
        self._lock = threading.Lock()
        self._check_existing_shards()

    # ------------------------------------------------------------------ resume
    def _check_existing_shards(self):
        for shard_name, names in self.shard_to_tensors.items():
            if self.only_shards is not None and shard_name not in self.only_shards:
                continue                                   # another rank may be writing it right now
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
            src = src.contiguous()
            buf = self._pool.take(src.numel() * src.element_size())
            host = buf[: src.numel() * src.element_size()].view(src.dtype).reshape(src.shape)
            produced = torch.cuda.current_stream(t.device).record_event()
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(produced)
                host.copy_(src, non_blocking=True)
                src.record_stream(self._copy_stream)
                done = self._copy_stream.record_event()
            self._events.setdefault(shard_name, []).append(done)
            self._recent_events.append(done)
            self._bufs.setdefault(shard_name, []).append(buf)
        else:
            host = t.clone().to(self.output_astype).contiguous()
        self._staged.setdefault(shard_name, {})[layer_name] = host

    def _flush(self, shard_name: str):
        staged = self._staged.pop(shard_name, None)
        if not staged:
            return
        events = self._events.pop(shard_name, [])
        bufs = self._bufs.pop(shard_name, [])

        def job():
            try:
                for ev in events:
                    ev.synchronize()
                self._write_shard(shard_name, staged)
            finally:
                for b in bufs:
                    self._pool.give(b)

        if self.async_write:
            if self._worker is None:
                self._worker = _ShardWriter(self.queue_depth, self.writer_threads, self.io_threads)
            self._worker.submit(job)
        else:
            job()

    def _write_shard(self, shard_name: str, staged: Dict[str, torch.Tensor]):
        path = self.output_path / shard_name
        tensors = dict(staged)
        if path.exists():                                   # resumed run: keep what an earlier run wrote
            with safe_open(path, framework="pt") as f:
                for key in f.keys():
                    tensors.setdefault(key, f.get_tensor(key))
        ordered = {n: tensors[n] for n in sorted(tensors, key=lambda n: self._order_pos.get(n, 1 << 30))}
        tmp = path.with_name(path.name + f".tmp{os.getpid()}")
        try:
            write_safetensors(tmp, ordered, metadata={"format": "pt"},
                              io_pool=self._worker.io_pool if self._worker is not None else None)
            os.replace(tmp, path)
            with self._lock:
                for name in staged:
                    self.written_shard_layers.add((shard_name, name))
            logger.info(f"Wrote shard {shard_name} ({len(ordered)} tensors)")
        except Exception as exc:                             # reference behaviour: log, drop the file (:146-149)
            logger.error(f"Error saving shard {shard_name}: {exc}")
            for p in (tmp, path):
                if p.exists():
                    p.unlink()

    # ------------------------------------------------------------------ reference API
    def add_tensor(self, layer_name: str, tensor: torch.Tensor):
        shard_name = self.base_index["weight_map"][layer_name]
        with self._lock:
            if (shard_name, layer_name) in self.written_shard_layers:
                logger.info(f"Skipping {layer_name} as it's already in written shard {shard_name}")
                return
            on_disk = {n for (s, n) in self.written_shard_layers if s == shard_name}
        self._stage(shard_name, layer_name, tensor)
        if on_disk | set(self._staged[shard_name]) >= self.shard_to_tensors[shard_name]:
            self._flush(shard_name)

    def wait_staged(self):
        """Block until every tensor handed to add_tensor so far has arrived in pinned host memory (the device -> host
        part of the hand-over; shard files may still be in the writer threads)."""
        events, self._recent_events = self._recent_events, []
        for ev in events:
            ev.synchronize()

    def wait(self):
        """Block until every shard handed to the writer thread is on disk."""
        if self._worker is not None:
            self._worker.drain()

    def flush_partial(self):
        """Write every staged tensor now, complete shard or not (a merge that aborts calls this so that the next run
        resumes after the last merged tensor, as with the reference's write-per-tensor)."""
        for shard_name in list(self._staged):
            self._flush(shard_name)
        if self._worker is not None:
            self._worker.drain()

    def finalize(self, only_shards=None):
        """Flush what is staged and verify completeness (shard/writer.py:151-161).  `only_shards`
        restricts the check to the shards this process owns (multi-GPU merge, schedule.py)."""
        self.flush_partial()
        if self._worker is not None:
            self._worker.close()
            errors, self._worker = self._worker.errors, None
            if errors:
                raise errors[0]
        only = only_shards if only_shards is not None else self.only_shards
        missing = [(s, n) for s, names in self.shard_to_tensors.items() for n in names
                   if (only is None or s in only) and (s, n) not in self.written_shard_layers]
        if missing:
            logger.error(f"Failed to write all layers. Missing: {missing}")
            raise RuntimeError(f"Incomplete model output: missing {len(missing)} layers")

    def shard_layers(self) -> Generator[List[ShardLayer], None, None]:
        for shard_name in sorted(self.shard_to_tensors):
            names = sorted(self.shard_to_tensors[shard_name], key=lambda n: self._order_pos[n])
            group = []
            for n in names:
                sl = ShardLayer(self._order_pos[n], shard_name, n, (shard_name, n) in self.written_shard_layers)
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
