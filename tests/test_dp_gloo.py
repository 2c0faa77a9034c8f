"""world_size-2 gloo tests of the data-parallel glue (host logic; NCCL on the GPU box)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lc2is_b200 import dp
from oracle import head_oracle as O


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w, _ = dp.init_distributed("gloo")
    assert (r, w) == (rank, world) and dp.is_dist()
    g = torch.Generator().manual_seed(11)
    N, C = 9, 6
    pred = torch.randint(0, C, (N, 8, 8), generator=g)
    lab = torch.randint(0, C, (N, 8, 8), generator=g)
    a, b = dp.shard_range(N, rank, world)
    # eval: local integer confusion matrix, one all-reduce
    cm = O.confusion_matrix(pred[a:b], lab[a:b], C)
    dp.allreduce_confmat_(cm)
    # per-image stats gather
    per = torch.stack([torch.stack([torch.diag(c), c.sum(1), c.sum(0)]) for c in
                       (O.confusion_matrix(pred[i], lab[i], C) for i in range(a, b))])
    counts = [dp.shard_range(N, k, world)[1] - dp.shard_range(N, k, world)[0] for k in range(world)]
    allper = dp.gather_per_image(per, counts)
    # train: N_valid-weighted mean.  local grads are scaled by 1/N_valid_global before the all-reduce.
    low = torch.randn(N, C, 2, 2, generator=g)
    labels = torch.randint(0, C, (N, 8, 8), generator=g)
    nv = (labels[a:b] != 0).sum().reshape(1)
    dp.global_valid_count_(nv)
    x = low[a:b].clone().requires_grad_(True)
    up = torch.nn.functional.interpolate(x, mode="bilinear", size=8)
    loss_sum = torch.nn.functional.cross_entropy(up, labels[a:b], ignore_index=0, reduction="sum")
    (loss_sum / nv).backward()
    w_shared = torch.ones(C, 2, 2)
    bucket = dp.GradBucket([(C, 2, 2), (1,)], device="cpu")
    bucket.views[0].copy_((x.grad * w_shared).sum(0))            # a "shared parameter" gradient
    bucket.views[1].copy_(loss_sum.detach().reshape(1))
    bucket.allreduce_()
    xa = torch.full((3,), float(rank + 1))
    wk = dp.allreduce_sum_async(xa)
    wk.wait()
    assert torch.equal(xa, torch.full((3,), 3.0))
    if rank == 0:
        q.put(dict(cm=cm, allper=allper, nv=nv, g=bucket.views[0].clone(), loss=bucket.views[1] / nv,
                   pred=pred, lab=lab, low=low, labels=labels))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    C = 6
    # all-reduced confusion matrix == single-process matrix on the concatenated set, bit for bit
    assert torch.equal(out["cm"], O.confusion_matrix(out["pred"], out["lab"], C))
    assert out["allper"].shape == (9, 3, C)
    for i in range(9):
        c = O.confusion_matrix(out["pred"][i], out["lab"][i], C)
        assert torch.equal(out["allper"][i, 0], torch.diag(c)) and torch.equal(out["allper"][i, 1], c.sum(1))
    # summed gradients == single-process gradient of the global-mean loss
    x = out["low"].clone().requires_grad_(True)
    up = torch.nn.functional.interpolate(x, mode="bilinear", size=8)
    loss = torch.nn.functional.cross_entropy(up, out["labels"], ignore_index=0)
    loss.backward()
    assert int(out["nv"]) == int((out["labels"] != 0).sum())
    torch.testing.assert_close(out["g"], x.grad.sum(0), rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(out["loss"].reshape(()), loss.detach(), rtol=1e-6, atol=1e-6)


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 2000):
        for w in (1, 2, 3, 8):
            rs = [dp.shard_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(w - 1))
            assert max(b - a for a, b in rs) - min(b - a for a, b in rs) <= 1
