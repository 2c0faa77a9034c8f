"""Ragged evaluation (compute_gt_mIOU shape): ONE lc2is_argmax_confmat_ragged launch vs one lc2is_argmax_confmat_lowres
(k3_low_gen_kernel) call per image.  usage: time_ragged.py [N]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lc2is_b200 import ops
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
C, h = 151, 128
g = torch.Generator().manual_seed(3)
sizes = [(int(torch.randint(384, 700, (1,), generator=g)), int(torch.randint(384, 700, (1,), generator=g))) for _ in range(N)]
low = (torch.randn(N, C, h, h, generator=g) * 0.05).cuda()
gts = [torch.randint(0, C, s, generator=g).cuda() for s in sizes]
flat = torch.cat([t.reshape(-1) for t in gts])
npx = sum(a * b for a, b in sizes)
def ragged():
    return ops.argmax_confmat_ragged(low, sizes, flat, mode="bicubic")
def loop():
    return [ops.argmax_confmat(low[i:i + 1], gts[i][None], per_image=True, size=sizes[i], mode="bicubic") for i in range(N)]
for name, fn in (("ragged (1 launch)", ragged), ("per image (k3_low_gen)", loop)):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name:28s} {ms:8.3f} ms for {N} images, {npx / 1e6:.1f} Mpx, C={C}: {npx * C / ms / 1e6:.1f} G class-pixels/s")
a = ragged()[1]; b = torch.cat([x[1] for x in loop()])
print("per-image stats equal:", bool(torch.equal(a, b)))
