import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lc2is_b200 import ops
M = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
x = torch.randn(M, 768, device="cuda").to(torch.bfloat16); w = (torch.randn(512, 768, device="cuda") * 0.036).to(torch.bfloat16)
b = torch.randn(512, device="cuda")
for _ in range(3): ops.linear_fwd(x, w, b, torch.bfloat16)
torch.cuda.synchronize(); print("ok")
