"""Engine shim (lc2is_b200/engine.py; reference engine.py:14-208): loop shape and data-parallel host logic on CPU
(gloo, world_size 2); the online evaluation and the CUDA-graph step on the GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

from lc2is_b200 import dp
from lc2is_b200.engine import Engine


class TinySeg(nn.Module):
    """inputs {'image': [B,3,h,w]} -> {'outputs': [B,C,h,w]} like the reference models' forward(inputs) -> dict."""

    def __init__(self, C=5):
        super().__init__()
        self.conv = nn.Conv2d(3, C, 1)

    def forward(self, inputs):
        return dict(outputs=self.conv(inputs["image"]))


def _batches(n, B, C, h, seed):
    g = torch.Generator().manual_seed(seed)
    return [(dict(image=torch.randn(B, 3, h, h, generator=g), label=torch.randint(0, C, (B, h, h), generator=g)), None)
            for _ in range(n)]


def _run(model, loader, eval_loader=None, **kw):
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    eng = Engine("t", model, opt, nn.CrossEntropyLoss(), device="cpu", train_loader=loader, eval_loader=eval_loader,
                 logger=None, save_step=10 ** 9, online_metrics=False, **kw)
    return eng, eng.train()


def test_engine_loop_shape_cpu(tmp_path):
    torch.manual_seed(0)
    model = TinySeg()
    loader = _batches(4, 2, 5, 8, 1)
    seen = {}

    def compute_metrics(outputs, labels):
        seen["shapes"] = (tuple(outputs.shape), tuple(labels.shape))
        return dict(acc=float((outputs.argmax(1) == labels).float().mean()))
    eng, (metrics, save_path) = _run(model, loader, eval_loader=_batches(3, 2, 5, 8, 2), compute_metrics=compute_metrics,
                                     eval_step=4, log_step=2, out_dir=str(tmp_path) + "/")
    assert eng.train_step == 4 and eng.stop_train
    assert seen["shapes"] == ((6, 5, 8, 8), (6, 8, 8))           # ONE concat of the three eval batches
    assert {"train_step", "train_epoch", "train_loss", "eval_loss", "eval_acc"} <= set(metrics)
    assert metrics["train_step"] == 4 and save_path is None


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q, tmp):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dp.init_distributed("gloo")
    torch.manual_seed(0)
    model = TinySeg()
    full = _batches(3, 4, 5, 8, 3)
    a, b = dp.shard_range(4, rank, world)
    shard = [(dict(image=d["image"][a:b], label=d["label"][a:b]), m) for d, m in full]
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    eng = Engine("dp", model, opt, nn.CrossEntropyLoss(), device="cpu", train_loader=shard, logger=None,
                 save_step=3, log_step=3, out_dir=tmp + "/", online_metrics=False)
    metrics, save_path = eng.train()
    if rank == 0:
        q.put(dict(state={k: v.tolist() for k, v in model.state_dict().items()}, metrics=metrics, save_path=save_path))
    dist.barrier()
    dist.destroy_process_group()


def test_engine_data_parallel_gloo_world2(tmp_path):
    """Two ranks on half batches == one process on the full batches (plain mean CE: the average of the shard gradients
    is the full-batch gradient); only rank 0 writes the checkpoint."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q, str(tmp_path))) for r in range(2)]
    for p in ps:
        p.start()
    got = q.get(timeout=120)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    ref = TinySeg()
    eng, (metrics, _) = _run(ref, _batches(3, 4, 5, 8, 3), log_step=3)
    for k, v in ref.state_dict().items():
        assert torch.allclose(torch.tensor(got["state"][k]), v, atol=1e-6), k
    assert abs(got["metrics"]["train_loss"] - metrics["train_loss"]) < 1e-6
    files = sorted(os.listdir(os.path.join(got["save_path"])))
    assert files == ["step-3.pt"]


@pytest.mark.gpu
def test_engine_online_eval_and_graph_step_gpu():
    """Head-shaped model built from the mirrors (TextToPatch projection -> cosine logits; AuxiliaryLoss on the low map):
    the online evaluation equals the metrics mirror on the concatenated outputs, and the CUDA-graph step trains like the
    eager step."""
    from lc2is_b200 import head, metrics
    from lc2is_b200.model.loss import AuxiliaryLoss
    dev = torch.device("cuda")
    C, D, h = 12, 64, 8

    class HeadModel(nn.Module):
        def __init__(self):
            super().__init__()
            self.visual = nn.Linear(32, D)
            self.t = nn.Parameter(torch.randn(C, D))

        def forward(self, inputs):
            v = self.visual(inputs["patches"])
            low = head.cosine_logits(v, self.t, hw_shape=(h, h))
            return dict(outputs=low, low_score_map=low)

    def batches(n, seed):
        g = torch.Generator().manual_seed(seed)
        return [(dict(patches=torch.randn(2, h * h, 32, generator=g), label=torch.randint(0, C, (2, 16 * h, 16 * h), generator=g)), None)
                for _ in range(n)]

    def train(graph):
        torch.manual_seed(1)
        m = HeadModel()
        opt = torch.optim.SGD(m.parameters(), lr=0.5)
        crit = AuxiliaryLoss(ignore_index=0)
        eng = Engine("g", m, opt, crit, aux_criterion=crit, device=dev, train_loader=batches(6, 4), logger=None,
                     save_step=10 ** 9, log_step=6, cuda_graph=graph)
        metrics_, _ = eng.train()
        return m, metrics_, eng
    m_e, met_e, _ = train(False)
    m_g, met_g, eng = train(True)
    assert eng._graph is not None and "graph" in eng._graph
    for (k, a), (_, b) in zip(m_e.state_dict().items(), m_g.state_dict().items()):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-5), k
    assert abs(met_e["train_loss"] - met_g["train_loss"]) < 1e-4

    # online evaluation: labels on the outputs' grid (metrics.py:84-90)
    g = torch.Generator().manual_seed(9)
    ev = [(dict(patches=torch.randn(2, h * h, 32, generator=g), label=torch.randint(0, C, (2, h, h), generator=g)), None)
          for _ in range(3)]

    class EvalCrit(nn.Module):
        def forward(self, out, lab):
            return nn.functional.cross_entropy(out, lab)
    eng = Engine("e", m_e, None, EvalCrit(), aux_criterion=EvalCrit(), device=dev, eval_loader=ev, logger=None,
                 ignore_index=0)
    got = eng.evaluate()
    with torch.no_grad():
        outs = torch.cat([m_e({"patches": d["patches"].to(dev)})["outputs"] for d, _ in ev]).cpu()
    labs = torch.cat([d["label"] for d, _ in ev])
    ref = metrics.compute_mIOU(outs, labs, n_cls=C, ignore_index=0)["mIOU_label"]
    assert abs(got["eval_mIOU_label"] - ref) < 1e-6
    up = torch.nn.functional.interpolate(outs, mode="bicubic", scale_factor=4)
    labu = labs.repeat_interleave(4, 1).repeat_interleave(4, 2)
    assert abs(got["eval_mIOU_global"] - metrics.compute_mIOU_tensor(up, labu, C, 0)) < 1e-5
    assert int(eng.last_eval_stats["confmat"].sum()) == labu.numel()
