import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
from lc2is_b200 import synthetic, dp, _lib
from lc2is_b200.step import HostStep
rank, world, lr = dp.init_distributed()
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
B, h, H, C = 16, 32, 512, 150
hv = synthetic.make_patch_embeddings(B, h * h, 512, seed=rank).pin_memory()
hl = synthetic.make_labels(B, H, H, C, ignore_frac=0.1, seed=rank).pin_memory()
t = synthetic.make_prototypes(C, 512).pin_memory()
cm_dev = torch.zeros(C, C, dtype=torch.int64, device=dev)
def bench(hp, coll, N=40):
    hs = HostStep(B, h, h, H, H, C, ignore_index=0, depth=2, host_pack=hp, device=dev)
    def fin(out):
        if coll:
            cm_dev.copy_(out[2], non_blocking=True); dist.all_reduce(cm_dev)
    for _ in range(3): hs(hv, t, hl)
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    hs.submit(hv, t, hl)
    for i in range(1, N):
        hs.submit(hv, t, hl); fin(hs.wait())
    fin(hs.wait())
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / N
    if rank == 0: print(f"world {world} threads {_lib.lib.lc2is_pack_threads()} host_pack {hp} collective {coll}: {dt*1e3:.3f} ms/step", flush=True)
for hp in (True, False):
    for coll in (False, True):
        bench(hp, coll)
if world > 1: dist.destroy_process_group()
