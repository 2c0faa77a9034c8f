"""Time the ContrastiveLoss kernels (K4) with CUDA events; algorithmic bytes: forward = outputs once + labels,
backward = outputs + gradient.  python tools/time_k4.py"""
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lc2is_b200 import ops

dev = "cuda"
res = []
cfgs = ((16, 32), (16, 128), (8, 128)) if len(sys.argv) < 3 else ((int(sys.argv[1]), int(sys.argv[2])),)
for B, h in cfgs:
    C = 151
    n_sets = max(2, int(300e6 // (B * h * h * C * 4)) + 1)          # rotate over > L2 worth of inputs
    outs = [torch.randn(B, h * h, C, device=dev) for _ in range(n_sets)]
    lab = torch.randint(0, C, (B, h, h), device=dev)
    coef = torch.tensor([0.5 / (B * h * h), 0.5 / (B * h * C)], device=dev)
    for _ in range(3):
        s, c, l, n = ops.contrastive_fwd(outs[0], lab, -100); ops.contrastive_bwd(outs[0], lab, -100, l, n, coef)
    torch.cuda.synchronize()
    it = 20
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tfs, tbs = [], []
    for i in range(it):
        o = outs[i % n_sets]
        e[0].record(); s, c, l, n = ops.contrastive_fwd(o, lab, -100)
        e[1].record(); g = ops.contrastive_bwd(o, lab, -100, l, n, coef)
        e[2].record(); torch.cuda.synchronize()
        tfs.append(e[0].elapsed_time(e[1])); tbs.append(e[1].elapsed_time(e[2]))
    tf, tb = sorted(tfs)[it // 2] * 1e3, sorted(tbs)[it // 2] * 1e3
    bytes_f = B * h * h * C * 4 + B * h * h * 8
    bytes_b = 2 * B * h * h * C * 4 + B * h * h * 8
    res.append(dict(B=B, h=h, fwd_us=round(tf, 1), bwd_us=round(tb, 1), fwd_GBs=round(bytes_f / tf / 1e3, 1),
                    bwd_GBs=round(bytes_b / tb / 1e3, 1)))
    print(res[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/k4_times.json", "w"))
