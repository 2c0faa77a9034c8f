"""Mirror of the reference ``metrics.py`` (same function names, signatures and result keys).

The per-image Python loop of the reference (bicubic x4 -> Softmax2d -> torchmetrics JaccardIndex,
metrics.py:87-99) is replaced by K3: one fused kernel that resizes in registers, takes the argmax
and accumulates the integer confusion matrix (global [C,C] and per-image TP / target / prediction
counts).  IoU itself is derived from those integers with the torchmetrics 0.10/0.11 formulas
(un-vendored dependency; restated in oracle/head_oracle.py, "parity unpinned"):
    iou_k = cm[k,k] / (cm[k,:].sum() + cm[:,k].sum() - cm[k,k]),  0.0 where the union is empty.

Inputs may live on the CPU (``Engine.eval_loop`` concatenates ``.cpu()`` tensors, engine.py:162);
they are streamed to the GPU in chunks.  There is no CPU implementation here.
"""
from __future__ import annotations

from typing import List, Optional

import torch
from torch import Tensor

from . import ops

_CHUNK_BYTES = 1 << 30


def _device() -> torch.device:
    if not torch.cuda.is_available():
        from ._lib import Lc2isError
        raise Lc2isError("lc2is_b200.metrics needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev(t: Tensor, dev, dtype=None) -> Tensor:
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.to(dev, non_blocking=True).contiguous()


def _iou_from_counts(tp: Tensor, row: Tensor, col: Tensor) -> Tensor:
    """torchmetrics `_jaccard_from_confmat` (average='none', absent_score=0): float32 division."""
    union = row + col - tp
    iou = tp.float() / union.float()
    return torch.where(union == 0, torch.zeros_like(iou), iou)


def _per_image_miou(per_image: Tensor, ignore_index: Optional[int]) -> Tensor:
    """metrics.py:91-97: mean IoU over the classes present in the label, minus ignore_index."""
    tp, row, col = per_image[:, 0], per_image[:, 1], per_image[:, 2]
    iou = _iou_from_counts(tp, row, col)
    present = row > 0                                    # label.unique()
    if ignore_index is not None and 0 <= ignore_index < present.shape[1]:
        present[:, ignore_index] = False
    cnt = present.sum(1)
    return (iou * present).sum(1) / cnt.float()          # empty -> nan, like torch.mean of nothing


def prepare_for_label_metrics(outputs: Tensor, labels: Tensor, scale_factor: int = 4):
    """metrics.py:25-33, kept for API parity (materialises; prefer compute_mIOU)."""
    import torch.nn.functional as F
    outputs = F.interpolate(input=outputs, mode="bicubic", scale_factor=scale_factor)
    labels = F.interpolate(input=labels.view(-1, 1, labels.shape[-1], labels.shape[-1]).float(), mode="nearest",
                           scale_factor=scale_factor).squeeze().long()
    return [x for x in outputs], [x for x in labels]


def prepare_for_gt_metrics(outputs: Tensor, gt_list: List[Tensor], sizes: Tensor):
    """metrics.py:35-42, kept for API parity."""
    import torch.nn.functional as F
    outputs_list = [F.interpolate(input=o.unsqueeze(0), mode="bicubic", size=tuple(int(x) for x in s)).squeeze()
                    for o, s in zip(outputs, sizes)]
    return outputs_list, gt_list


def compute_mIOU(outputs: Tensor, labels: Tensor, n_cls: int, ignore_index: Optional[int] = 0) -> dict:
    """metrics.py:82-102.  outputs [N,C,h,w] float, labels [N,h,w] int64 (same grid as outputs; both are
    upsampled x4 - bicubic / nearest - inside the kernel)."""
    dev = _device()
    N, C, h, w = outputs.shape
    if C != n_cls:
        raise ValueError(f"outputs have {C} classes, n_cls={n_cls}")
    per_img_bytes = C * h * w * 4
    step = max(1, _CHUNK_BYTES // per_img_bytes)
    mious = []
    for i in range(0, N, step):
        o = _to_dev(outputs[i:i + step], dev, torch.float32)
        l = _to_dev(labels[i:i + step], dev, torch.int64)
        _, pi, _ = ops.argmax_confmat(o, l, per_image=True, size=(4 * h, 4 * w), mode="bicubic")
        mious.append(_per_image_miou(pi, ignore_index))
    return dict(mIOU_label=torch.concat(mious).mean().item())


def compute_gt_mIOU(outputs: Tensor, gt_list: List[Tensor], sizes: Tensor, n_cls: int = 151,
                    ignore_index: Optional[int] = 0) -> dict:
    """metrics.py:61-79: bicubic to each image's ORIGINAL size against ragged ground truth.  ONE launch per chunk of
    images (lc2is_argmax_confmat_ragged over a descriptor table) instead of one resize + metric call per image."""
    dev = _device()
    N = len(gt_list)
    C, h, w = outputs.shape[1:]
    if C != n_cls:
        raise ValueError(f"outputs have {C} classes, n_cls={n_cls}")
    step = max(1, _CHUNK_BYTES // (C * h * w * 4))
    mious = []
    for i in range(0, N, step):
        j = min(N, i + step)
        sz = [(int(sizes[k][0]), int(sizes[k][1])) for k in range(i, j)]
        flat = torch.cat([gt_list[k].reshape(-1).to(torch.int64) for k in range(i, j)])
        _, pi, _, _ = ops.argmax_confmat_ragged(_to_dev(outputs[i:j], dev, torch.float32), sz,
                                                _to_dev(flat, dev, torch.int64), mode="bicubic")
        mious.append(_per_image_miou(pi, ignore_index))
    return dict(mIOU_gt=torch.concat(mious).mean().item())


def confusion_matrix(pred: Tensor, label: Tensor, confmat: Optional[Tensor] = None) -> Tensor:
    """Global int64 [C,C] confusion matrix (rows = target) of argmax(pred) - the integer the data-parallel
    eval all-reduces.  pred [N,C,H,W], label [N,H,W]."""
    dev = _device()
    N, C, H, W = pred.shape
    per_img_bytes = C * H * W * pred.element_size()
    step = max(1, _CHUNK_BYTES // per_img_bytes)
    for i in range(0, N, step):
        p = pred[i:i + step]
        if p.dtype not in (torch.float32, torch.bfloat16):
            p = p.float()
        confmat, _, _ = ops.argmax_confmat(_to_dev(p, dev), _to_dev(label[i:i + step], dev, torch.int64), confmat=confmat)
    return confmat


def miou_from_confmat(cm: Tensor, ignore_index: Optional[int] = 0) -> Tensor:
    """``JaccardIndex(num_classes, ignore_index)`` (macro) from a confusion matrix: zero the ignore row,
    per-class IoU with absent classes = 0.0, drop the ignore class, mean (metrics.py:130)."""
    cm = cm.clone()
    C = cm.shape[0]
    keep = torch.ones(C, dtype=torch.bool, device=cm.device)
    if ignore_index is not None and 0 <= ignore_index < C:
        cm[ignore_index] = 0
        keep[ignore_index] = False
    iou = _iou_from_counts(torch.diag(cm), cm.sum(1), cm.sum(0))
    return iou[keep].mean()


def pixel_accuracy_from_confmat(cm: Tensor, ignore_index: Optional[int] = 0) -> Tensor:
    """Not in the reference: trace / sum over the non-ignored target rows."""
    cm = cm.clone()
    if ignore_index is not None and 0 <= ignore_index < cm.shape[0]:
        cm[ignore_index] = 0
    return torch.diag(cm).sum().double() / cm.sum().double()


def compute_mIOU_tensor(pred: Tensor, label: Tensor, n_cls: int, ignore_index: Optional[int] = 0) -> float:
    """metrics.py:127-134: dataset-level JaccardIndex over the whole tensor."""
    if pred.shape[1] != n_cls:
        raise ValueError(f"pred has {pred.shape[1]} classes, n_cls={n_cls}")
    return miou_from_confmat(confusion_matrix(pred, label), ignore_index).item()


def segmentation_metrics(outputs: Tensor, labels: Tensor, gt_list: List[Tensor], sizes: Tensor, n_clas: int = 151,
                         ignore_index: Optional[int] = 0) -> dict:
    """metrics.py:45-58."""
    segm_metrics = {}
    segm_metrics.update(compute_mIOU(outputs=outputs, labels=labels, n_cls=n_clas, ignore_index=ignore_index))
    segm_metrics.update(compute_gt_mIOU(outputs=outputs, gt_list=gt_list, sizes=sizes, n_cls=n_clas,
                                        ignore_index=ignore_index))
    return segm_metrics


def original_size_interpolate(tensor: Tensor, ori_size: Tensor) -> List[Tensor]:
    """metrics.py:137-143 (API parity; materialises with torch)."""
    import torch.nn.functional as F
    return [F.interpolate(input=t.unsqueeze(0), mode="bicubic", size=tuple(int(x) for x in s)).squeeze()
            for t, s in zip(tensor, ori_size)]


def pad_and_concat(tensor_list: List[Tensor], ori_size: Tensor, pad: str = "max", value: int = 0) -> Tensor:
    """metrics.py:145-157."""
    import torch.nn.functional as F
    max_size = ori_size.max(0).values if pad == "max" else torch.LongTensor([1024, 1024])
    padded = [F.pad(n, pad=(0, int(max_size[1] - s[1]), 0, int(max_size[0] - s[0])), mode="constant",
                    value=value).unsqueeze(0) for n, s in zip(tensor_list, ori_size)]
    return torch.cat(padded, dim=0)


def unpad(tensor: Tensor, size: Tensor) -> List[Tensor]:
    """metrics.py:159-165."""
    return [t[: int(s[0]), : int(s[1])] for t, s in zip(tensor, size)]


def reshape_tensor(tensor: Tensor, ori_size: Tensor) -> Tensor:
    """metrics.py:167-172."""
    return pad_and_concat(original_size_interpolate(tensor, ori_size), ori_size, pad="max", value=0)
