"""model.py:42-44 at the G-B size (B=16, 32x32 tokens of 768 channels -> 128x128): the fused kernel vs the reference lines
run by torch on the same GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from lc2is_b200 import ops
B, h, C = 16, 32, 768
x = torch.randn(B, h * h, C, device="cuda")
gy = torch.randn(B, 16 * h * h, C, device="cuda").to(torch.bfloat16)
def ref(x):
    t = x.permute(0, 2, 1).reshape(B, C, h, h)
    t = F.interpolate(t, mode="bicubic", scale_factor=4)
    return t.reshape(B, C, 16 * h * h).permute(0, 2, 1).contiguous().to(torch.bfloat16)
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
out_mb = B * 16 * h * h * C * 2 / 1e6
us = t(lambda: ops.bicubic4_tokens_fwd(x, (h, h))); print(f"ours  fwd {us:8.1f} us  ({out_mb / us:.2f} TB/s of output)")
us = t(lambda: ref(x)); print(f"torch fwd {us:8.1f} us")
us = t(lambda: ops.bicubic4_tokens_bwd(gy, (h, h))); print(f"ours  bwd {us:8.1f} us")
xr = x.clone().requires_grad_(True)
def tb():
    y = ref(xr); y.backward(gy)
us = t(tb); print(f"torch fwd+bwd {us:8.1f} us")
