import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lc2is_b200 import _lib
lib = _lib.lib
lab = torch.randint(0, 150, (16, 512, 512)).pin_memory()
out = torch.empty(16, 512, 512, dtype=torch.uint16).pin_memory()
for _ in range(3): lib.lc2is_pack_labels_host(lab.data_ptr(), lab.numel(), 150, 0, out.data_ptr())
t0 = time.perf_counter()
for _ in range(20): lib.lc2is_pack_labels_host(lab.data_ptr(), lab.numel(), 150, 0, out.data_ptr())
dt = (time.perf_counter() - t0) / 20
print("threads", os.environ.get("LC2IS_PACK_THREADS", "default"), "pack ms %.3f  GB/s %.1f" % (dt * 1e3, lab.numel() * 8 / dt / 1e9))
