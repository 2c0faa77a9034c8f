import torch
n = 24 * 2**20
x = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
def run(piece, reps=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for o in range(0, n, piece): d[o:o+piece].copy_(x[o:o+piece], non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    return n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
for _ in range(2): run(n)
for piece in (n, 12 * 2**20, 8 * 2**20, 4 * 2**20, 2 * 2**20, 1 * 2**20):
    print(f"24 MiB as pieces of {piece/2**20:.0f} MiB: {run(piece):.1f} GB/s")
