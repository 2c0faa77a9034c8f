import cProfile, pstats, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lc2is_b200 import synthetic
from lc2is_b200.step import HeadStep
B, h, H, C = 16, 32, 512, 150
dev = "cuda"
v = synthetic.make_patch_embeddings(B, h * h, 512).to(dev); t = synthetic.make_prototypes(C, 512).to(dev)
l = synthetic.make_labels(B, H, H, C, ignore_frac=0.1).to(dev)
step = HeadStep(B, h, h, H, H, C, ignore_index=0)
for _ in range(20): step(v, t, l)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step(v, t, l)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
