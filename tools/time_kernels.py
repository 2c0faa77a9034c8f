"""Per-kernel CUDA-event timings at the BASELINE geometries (development aid; not the bench).
Inputs rotate over several buffers larger than L2 in total."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lc2is_b200 import ops, synthetic, _lib

def timeit(fn, n=30, warm=5):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(warm + i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3   # us

def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--geom", default="A"); ap.add_argument("--B", type=int, default=16)
    ap.add_argument("--C", type=int, default=150); ap.add_argument("--only", default="")
    a = ap.parse_args()
    h = {"A": 32, "B": 128, "E": 64}[a.geom]; H = 1024 if a.geom == "E" else 512
    B, C, D = a.B, a.C, 512
    dev = "cuda"
    nset = 4 if a.geom == "A" else 2
    V = [synthetic.make_patch_embeddings(B, h * h, D, seed=i).to(dev) for i in range(nset)]
    L = [synthetic.make_labels(B, H, H, C, seed=i, ignore_frac=0.1).to(dev) for i in range(nset)]
    t = synthetic.make_prototypes(C, D).to(dev)
    t_hat, inv_t = ops.proto_normalize(t)
    res = {}
    outs = [ops.cosine_logits_fwd(V[i], t_hat, C, (h, h)) for i in range(nset)]
    def want(k): return not a.only or k in a.only.split(",")
    if want("k1"):
        res["k1_fwd(prep+gemm)"] = timeit(lambda i: ops.cosine_logits_fwd(V[i % nset], t_hat, C, (h, h)))
    nv = ops.count_valid(L[0], C, 0); gs = ops.mean_scale(nv)
    if want("count"):
        res["count_valid"] = timeit(lambda i: ops.count_valid(L[i % nset], C, 0))
    if want("k2"):
        res["k2(memset+fused)"] = timeit(lambda i: ops.upsample_ce(outs[i % nset][0], L[i % nset], 0, gs))
        if ops.ce_split_supported(h, h, H, H):
            res["k2_split(memset+prepass+strip)"] = timeit(lambda i: ops.upsample_ce_split(outs[i % nset][0], L[i % nset], 0))
        res["k2_fwd_only"] = timeit(lambda i: ops.upsample_ce(outs[i % nset][0], L[i % nset], 0, gs, want_grad=False))
    _, g, gb = ops.upsample_ce(outs[0][0], L[0], 0, gs, want_bf16=True)
    if want("k1b"):
        res["grad_to_bf16"] = timeit(lambda i: ops.grad_to_bf16(g))
        res["k1b(prep+dv+dt+finish)"] = timeit(lambda i: ops.cosine_logits_bwd(gb, outs[i % nset][0], outs[i % nset][1], outs[i % nset][2], t_hat, inv_t, C, grad_v_dtype=torch.bfloat16))
    if want("k3"):
        res["k3_lowres_bilinear"] = timeit(lambda i: ops.argmax_confmat(outs[i % nset][0], L[i % nset], size=(H, H), mode="bilinear"))
        if ops.ce_split_supported(h, h, H, H):
            pk = ops.upsample_ce_split(outs[0][0], L[0], 0, want_grad=False)[3]
            res["k3_strip_packed"] = timeit(lambda i: ops.argmax_confmat_packed(outs[i % nset][0], pk, (H, H)))
        res["k3_lowres_bicubic"] = timeit(lambda i: ops.argmax_confmat(outs[i % nset][0], L[i % nset], size=(H, H), mode="bicubic"))
    if want("k3full"):
        n3 = 4
        full = [torch.randn(n3, C, H, H, device=dev) for _ in range(2)]
        res["k3_full_fp32_per_img"] = timeit(lambda i: ops.argmax_confmat(full[i % 2], L[i % nset][:n3]), n=10) / n3
        by = n3 * C * H * H * 4 + n3 * H * H * 8
        res["k3_full_GBps"] = by / (res["k3_full_fp32_per_img"] * n3 * 1e-6) / 1e9
        del full
    print(json.dumps({"geom": a.geom, "B": B, "C": C, "us": {k: round(v, 1) for k, v in res.items()}}))

if __name__ == "__main__":
    main()
