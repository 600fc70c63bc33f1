"""development: can D2H copies land directly in a mapped tmpfs file?  cudaHostRegister on an mmap of /dev/shm:
registration cost (with / without parallel pre-faulting), copy bandwidth into it, is_pinned()."""
import mmap, os, sys, time
from concurrent.futures import ThreadPoolExecutor
import torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 30)
dev = torch.device("cuda:0")
src = torch.empty(n, dtype=torch.uint8, device=dev).fill_(7)
rt = torch.cuda.cudart()
pool = ThreadPoolExecutor(16)
for prefault in (0, 1):
    path = f"/dev/shm/probe_{os.getpid()}_{prefault}"
    fd = os.open(path, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o644)
    os.ftruncate(fd, n)
    mm = mmap.mmap(fd, n)
    t = torch.frombuffer(mm, dtype=torch.uint8)
    t0 = time.perf_counter()
    if prefault:
        ch = 32 << 20
        list(pool.map(lambda a: t[a:a + ch].zero_(), range(0, n, ch)))
    t1 = time.perf_counter()
    rc = rt.cudaHostRegister(t.data_ptr(), n, 0)
    t2 = time.perf_counter()
    print(f"prefault={prefault}: prefault {t1 - t0:.3f} s, cudaHostRegister rc={rc} {t2 - t1:.3f} s ({n / (t2 - t0) / 1e9:.2f} GB/s overall), is_pinned={t.is_pinned()}")
    if int(rc) == 0:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); t.copy_(src, non_blocking=True); e1.record(); torch.cuda.synchronize()
        print(f"   D2H into the mapping: {n / e0.elapsed_time(e1) / 1e6:.1f} GB/s; data ok: {bool((t[::4096] == 7).all())}")
        t3 = time.perf_counter(); rc2 = rt.cudaHostUnregister(t.data_ptr()); t4 = time.perf_counter()
        print(f"   cudaHostUnregister rc={rc2} {t4 - t3:.3f} s")
    del t
    mm.close(); os.close(fd)
    with open(path, "rb") as fh:
        fh.seek(n // 2); print("   file byte:", fh.read(1))
    os.unlink(path)
# pinned pool buffer reference
p = torch.empty(n, dtype=torch.uint8).pin_memory()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); p.copy_(src, non_blocking=True); e1.record(); torch.cuda.synchronize()
print(f"D2H into torch pinned memory: {n / e0.elapsed_time(e1) / 1e6:.1f} GB/s")
