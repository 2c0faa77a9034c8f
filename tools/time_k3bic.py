"""compute_mIOU's kernel: bicubic x4 + argmax + per-image counts at [16,151,128,128] -> 512^2 (k3_low_fast_kernel<1,1>)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lc2is_b200 import ops, synthetic
N, C, h = 16, 151, 128
low = (torch.randn(N, C, h, h, device="cuda") * 0.05)
lab = torch.randint(0, C, (N, h, h), device="cuda")
fn = lambda: ops.argmax_confmat(low, lab, per_image=True, size=(4 * h, 4 * h), mode="bicubic")
fn(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): fn()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 10 * 1e3
print(f"bicubic x4 argmax+confmat: {us:.1f} us for {N} images ({N * C * 512 * 512 / us / 1e6:.2f} T upsampled elements/s)")
