"""TextToPatch.visual backward at the G-B size (M = 16 x 128^2): lc2is_linear_bwd vs torch (cuBLAS) on the same bf16 operands."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lc2is_b200 import ops
M = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
K, N = 768, 512
x = torch.randn(M, K, device="cuda").to(torch.bfloat16); w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
gy = torch.randn(M, N, device="cuda").to(torch.bfloat16)
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
fl = 2.0 * M * K * N
for name, fn in (("ours  dX", lambda: ops.linear_bwd(gy, x, w, need_gw=False, need_gb=False)),
                 ("torch dX", lambda: gy @ w),
                 ("ours  dW (fp32 out)", lambda: ops.linear_bwd(gy, x, w, need_gx=False, need_gb=False)),
                 ("torch dW (bf16 out)", lambda: gy.t() @ x),
                 ("ours  db", lambda: ops.linear_bwd(gy, x, w, need_gx=False, need_gw=False)),
                 ("torch db", lambda: gy.float().sum(0)),
                 ("ours  all", lambda: ops.linear_bwd(gy, x, w))):
    us = t(fn)
    print(f"{name:22s} {us:9.1f} us  {fl / us / 1e6:8.1f} TFLOP/s-equivalent")
