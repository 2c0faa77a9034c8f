"""TextToPatch.visual projection: lc2is_linear_fwd vs torch (cuBLAS) bf16, TFLOP/s (development aid)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lc2is_b200 import ops
def t(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
K, N = 768, 512
for M in (16384, 65536, 262144):
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16); w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    b = torch.randn(N, device="cuda")
    us = t(lambda: ops.linear_fwd(x, w, b, torch.bfloat16))
    us_t = t(lambda: torch.nn.functional.linear(x, w, b.to(torch.bfloat16)))
    fl = 2.0 * M * N * K
    print(f"M={M}: lc2is_linear_fwd {us:.1f} us = {fl/us/1e6:.0f} TFLOP/s | torch F.linear bf16 {us_t:.1f} us = {fl/us_t/1e6:.0f} TFLOP/s")
