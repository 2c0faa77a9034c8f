"""Run one kernel a few times (for ncu captures).  usage: run_one.py {k2,k3low,k3bic,k3full,k1,k1b} [geom] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lc2is_b200 import ops, synthetic
which = sys.argv[1]; geom = sys.argv[2] if len(sys.argv) > 2 else "A"; B = int(sys.argv[3]) if len(sys.argv) > 3 else 4
h = {"A": 32, "B": 128}[geom]; H = 512; C = 150; D = 512; dev = "cuda"
V = synthetic.make_patch_embeddings(B, h * h, D).to(dev)
L = synthetic.make_labels(B, H, H, C, ignore_frac=0.1).to(dev)
t = synthetic.make_prototypes(C, D).to(dev)
t_hat, inv_t = ops.proto_normalize(t)
logits, v_hat, inv_v = ops.cosine_logits_fwd(V, t_hat, C, (h, h))
nv = ops.count_valid(L, C, 0); gs = ops.mean_scale(nv)
for _ in range(3):
    if which == "k1": ops.cosine_logits_fwd(V, t_hat, C, (h, h))
    if which == "k2": ops.upsample_ce(logits, L, 0, gs)
    if which == "k2split": ops.upsample_ce_split(logits, L, 0)
    if which == "step":
        from lc2is_b200.step import HeadStep
        if "hs" not in globals():
            hs = HeadStep(B, h, h, H, H, C, ignore_index=0)
        hs(V, t, L)
    if which == "k3split":
        pk = ops.upsample_ce_split(logits, L, 0, want_grad=False)[3]
        ops.argmax_confmat_packed(logits, pk, (H, H))
    if which == "k3low": ops.argmax_confmat(logits, L, size=(H, H), mode="bilinear")
    if which == "k3bic": ops.argmax_confmat(logits, L, size=(H, H), mode="bicubic")
    if which == "k3full":
        full = torch.randn(B, C, H, H, device=dev); ops.argmax_confmat(full, L)
    if which == "k1b":
        _, g, gb = ops.upsample_ce(logits, L, 0, gs, want_bf16=True)
        ops.cosine_logits_bwd(gb, logits, v_hat, inv_v, t_hat, inv_t, C, grad_v_dtype=torch.bfloat16)
torch.cuda.synchronize()
print("ok")
