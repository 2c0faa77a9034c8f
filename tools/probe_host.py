"""Host / PCIe probe of the GPU box (development aid)."""
import os, time, torch
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
os.system("lscpu | egrep 'Model name|Socket|Core|Thread|MHz|L3|NUMA' | head -12")
x = torch.empty(64 * 2**20, dtype=torch.uint8).pin_memory()
d = torch.empty_like(x, device="cuda")
for nbytes in (1 << 20, 8 << 20, 32 << 20, 64 << 20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): d[:nbytes].copy_(x[:nbytes], non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(f"H2D {nbytes>>20} MiB: {nbytes*10/ (e0.elapsed_time(e1)*1e-3)/1e9:.1f} GB/s")
# CPU int64->u16 pack rate, single thread numpy
import numpy as np
a = np.random.randint(0, 150, 4 * 2**20).astype(np.int64)
t0 = time.perf_counter()
for _ in range(5): b = a.astype(np.uint16)
dt = (time.perf_counter() - t0) / 5
print(f"numpy int64->uint16 1 thread: {a.nbytes/dt/1e9:.1f} GB/s read")
t0 = time.perf_counter()
for _ in range(5): c = a.copy()
dt = (time.perf_counter() - t0) / 5
print(f"numpy memcpy 32MB 1 thread: {a.nbytes/dt/1e9:.1f} GB/s")
