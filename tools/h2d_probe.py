import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lc2is_b200 import _lib
lib = _lib.lib
n = 16 * 512 * 512
lab = torch.randint(0, 150, (n,)).pin_memory()
out = torch.empty(n, dtype=torch.uint16).pin_memory()
d = torch.empty(n, dtype=torch.uint16, device="cuda")
def h2d():
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); d.copy_(out, non_blocking=True); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3
for _ in range(3): h2d()
print("H2D 8.4MB untouched: %.0f us" % h2d())
lib.lc2is_pack_labels_host(lab.data_ptr(), n, 150, 0, out.data_ptr())
print("H2D right after pack (lib, NT stores): %.0f us" % h2d())
print("H2D again: %.0f us" % h2d())
out.copy_(torch.randint(0, 150, (n,)).to(torch.uint16))
print("H2D right after torch cpu write: %.0f us" % h2d())
print("H2D again: %.0f us" % h2d())
