import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lc2is_b200 import synthetic
from lc2is_b200.step import HostStep
B, h, H, C = 16, 32, 512, 150
hv = synthetic.make_patch_embeddings(B, h * h, 512).pin_memory()
hl = synthetic.make_labels(B, H, H, C, ignore_frac=0.1).pin_memory()
t = synthetic.make_prototypes(C, 512).pin_memory()
for hp in (True, False):
    hs = HostStep(B, h, h, H, H, C, ignore_index=0, host_pack=hp)
    for _ in range(5): hs(hv, t, hl)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(30): hs(hv, t, hl)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 30
    print("host_pack", hp, "chunks", os.environ.get("LC2IS_STEP_CHUNKS", "default"), "ms/step %.3f" % (dt * 1e3), flush=True)

hs = HostStep(B, h, h, H, H, C, ignore_index=0, depth=2)
for _ in range(3): hs(hv, t, hl)
ref = (float(hs.out_loss), int(hs.out_n_valid), hs.out_confmat.clone())
torch.cuda.synchronize(); t0 = time.perf_counter()
N = 40
hs.submit(hv, t, hl)
for i in range(1, N):
    hs.submit(hv, t, hl)
    l, nv, cm = hs.wait()
    assert int(nv) == ref[1] and torch.equal(cm, ref[2])
hs.wait()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / N
print("submit/wait depth 2: ms/step %.3f" % (dt * 1e3), "loss", float(hs.out_loss), ref[0])
