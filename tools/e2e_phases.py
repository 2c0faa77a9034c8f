"""Host-side phase timing of the submit / prefetch / wait loop (what bench.py's e2e runs): where does the step period
go?  python tools/e2e_phases.py [depth]"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lc2is_b200 import synthetic, dp, _lib
from lc2is_b200.step import HostStep
B, h, H, C = 16, 32, 512, 150
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = 0
if world > 1:                                   # torchrun: every rank runs the loop, rank 0 prints
    import torch.distributed as dist
    rank, world, lr = dp.init_distributed()
    torch.cuda.set_device(lr)
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 2
RAW = int(sys.argv[2]) if len(sys.argv) > 2 else None
NB = 6
hv = [synthetic.make_patch_embeddings(B, h * h, 512, seed=i + 100 * rank).pin_memory() for i in range(NB)]
hl = [synthetic.make_labels(B, H, H, C, ignore_frac=0.1, seed=i + 100 * rank).pin_memory() for i in range(NB)]
t = synthetic.make_prototypes(C, 512).pin_memory()
for use_prefetch in (True, False):
    hs = HostStep(B, h, h, H, H, C, ignore_index=0, depth=depth, raw_images=RAW)
    N = 60
    ph = {"submit": [], "prefetch": [], "wait": []}
    for d in range(depth - 1):
        hs.submit(hv[d % NB], t, hl[d % NB])
    if use_prefetch:
        hs.prefetch(hl[(depth - 1) % NB])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t00 = time.perf_counter()
    for i in range(depth - 1, N):
        a = time.perf_counter()
        hs.submit(hv[i % NB], t, hl[i % NB])
        b = time.perf_counter()
        if use_prefetch and i + 1 < N:
            hs.prefetch(hl[(i + 1) % NB])
        c = time.perf_counter()
        hs.wait()
        d = time.perf_counter()
        ph["submit"].append(b - a); ph["prefetch"].append(c - b); ph["wait"].append(d - c)
    while hs._inflight:
        hs.wait()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t00) / (N - depth + 1)
    med = {k: sorted(v)[len(v) // 2] * 1e3 for k, v in ph.items()}
    if rank == 0:
      print("world", world, "threads", _lib.lib.lc2is_pack_threads(), "n_raw", hs.n_raw, getattr(hs, "calibration", None), "depth", depth, "prefetch", use_prefetch, "ms/step %.3f" % (dt * 1e3),
          " ".join("%s %.3f" % kv for kv in med.items()), flush=True)
if world > 1:
    dist.destroy_process_group()
