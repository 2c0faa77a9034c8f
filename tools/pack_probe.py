"""Host label-packing throughput (int64 -> host label form) over buffers larger than the last-level cache.
LC2IS_PACK_THREADS=n python tools/pack_probe.py"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lc2is_b200 import _lib
lib = _lib.lib
NB = 8
labs = [torch.randint(0, 150, (16, 512, 512)).pin_memory() for _ in range(NB)]
out = torch.empty(16, 512, 512, dtype=torch.uint16).pin_memory()
for i in range(3): lib.lc2is_pack_labels_host(labs[i].data_ptr(), labs[i].numel(), 150, 0, out.data_ptr())
ts = []
for i in range(40):
    lab = labs[i % NB]
    t0 = time.perf_counter()
    lib.lc2is_pack_labels_host(lab.data_ptr(), lab.numel(), 150, 0, out.data_ptr())
    ts.append(time.perf_counter() - t0)
dt = sorted(ts)[len(ts) // 2]
print("threads", lib.lc2is_pack_threads(), "pack ms %.3f  GB/s %.1f" % (dt * 1e3, labs[0].numel() * 8 / dt / 1e9), flush=True)
