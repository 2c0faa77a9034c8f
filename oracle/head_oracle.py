"""CPU oracle for the LC2IS segmentation-head hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` leg may import this module.  Nothing under ``lc2is_b200/``
imports it: the product path has no CPU fallback.

Every function restates reference lines with *live* torch CPU ops (torch is the
reference's own arithmetic dependency and is installed here), citing
``/root/reference`` file:line.  It never reads /root/reference at run time.

Pinning status
--------------
* cosine logits / bilinear upsample / cross-entropy: pinned.  ``tests/golden/
  make_golden.py`` imports the reference's own ``model.loss.AuxiliaryLoss`` and
  ``model.text_patch.TextToPatch`` and restates ``model/final.py:41-44`` verbatim;
  ``tests/test_oracle_golden.py`` checks this oracle against those fixtures bit-for-bit.
* ``contrastive_loss``: pinned against the reference's own ``model.loss.ContrastiveLoss`` (fixture
  ``tests/golden/contrastive.pt``: three losses + autograd gradient; fp32 tolerance, the restatement sums in another
  order than ``nn.CrossEntropyLoss``).
* confusion matrix / IoU (``jaccard_*``): **parity unpinned**.  The reference delegates
  to ``torchmetrics.JaccardIndex`` (un-vendored, version unpinned: requirements.txt:1
  is a placeholder; the ``JaccardIndex(num_classes=..)`` call form without ``task=``
  implies torchmetrics 0.10-0.11).  torchmetrics is not installed and there is no
  network, so its published algorithm is restated:
  ``confmat = bincount(target*C + pred, minlength=C*C).reshape(C, C)`` (rows=target),
  ``confmat[ignore_index] = 0``, ``iou = diag / (rowsum + colsum - diag)``,
  absent classes (union == 0) score 0.0, the ignore class is dropped, macro mean.
  The reference holds no golden vector for this boundary.  The restated definitions are cross-checked against an
  independent implementation that is installed (scikit-learn ``confusion_matrix`` / ``jaccard_score``:
  ``tests/test_oracle_golden.py::test_iou_formulas_against_scikit_learn``) - not the same as pinning against torchmetrics.
* ``upsample_tokens_bicubic4`` (model/model.py:42-44): the reference lines themselves (einops rearranges as permutes).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import torch
import torch.nn.functional as F
from torch import Tensor


# --------------------------------------------------------------------------------------
# Stage 1: L2 normalise + cosine logits                      (model/final.py:41-43)
# --------------------------------------------------------------------------------------
def l2_normalize(x: Tensor, dim: int) -> Tensor:
    """``F.normalize(x, dim=dim, p=2)`` = x / max(||x||_2, 1e-12).  final.py:41-42."""
    return F.normalize(x, dim=dim, p=2)


def cosine_logits(v: Tensor, t: Tensor, normalize: bool = True, logit_scale: float = 1.0,
                  hw_shape: Optional[Sequence[int]] = None) -> Tensor:
    """Head score map.

    v: [B, P, D] patch embeddings (the reference rearranges 'b (h w) c -> b c h w',
       final.py:37), t: [K, D] shared or [B, K, D] per-image text embeddings
       (final.py:31 ``expand(B,-1,-1)`` / final.py:129-130).
    normalize=True  -> final.py:41-43: normalize(v, dim=1), normalize(t, dim=2),
                       einsum('bchw,bkc->bkhw').
    normalize=False -> model/model.py:50,53: matmul(feature_v, feature_t.T) + rearrange.
    logit_scale: the reference has no temperature (only commented out, model.py:70,92);
                 1.0 reproduces it.
    Returns [B, K, h, w].
    """
    B, P, D = v.shape
    if hw_shape is None:
        h = int(round(P ** 0.5))
        hw_shape = (h, P // h)
    h, w = hw_shape
    assert h * w == P
    vm = v.transpose(1, 2).reshape(B, D, h, w)                   # b (h w) c -> b c h w
    tt = t if t.dim() == 3 else t.unsqueeze(0).expand(B, -1, -1)  # final.py:31
    if normalize:
        vm = F.normalize(vm, dim=1, p=2)                          # final.py:41
        tt = F.normalize(tt, dim=2, p=2)                          # final.py:42
    score = torch.einsum('bchw,bkc->bkhw', vm, tt)                # final.py:43
    if logit_scale != 1.0:
        score = score * logit_scale
    return score


def cosine_logits_bf16_operands(v: Tensor, t: Tensor, normalize: bool = True,
                                logit_scale: float = 1.0, hw_shape=None) -> Tensor:
    """Same as :func:`cosine_logits` but with the operands the tensor-core GEMM sees:
    normalisation in fp32, then both operands rounded to bf16 and upcast, fp32 accumulate.
    (SURVEY 8c parity rule: "oracle fed the same bf16-rounded operands".)"""
    B, P, D = v.shape
    v32 = v.float()
    t32 = t.float()
    if normalize:
        v32 = F.normalize(v32, dim=2, p=2)
        t32 = F.normalize(t32, dim=-1, p=2)
    v32 = v32.to(torch.bfloat16).float()
    t32 = t32.to(torch.bfloat16).float()
    return cosine_logits(v32, t32, normalize=False, logit_scale=logit_scale, hw_shape=hw_shape)


# --------------------------------------------------------------------------------------
# Stage 2: bilinear upsample + softmax cross-entropy        (model/loss.py:12-21)
# --------------------------------------------------------------------------------------
def upsample_bilinear(low: Tensor, size: Optional[int] = None, scale_factor: Optional[int] = None) -> Tensor:
    """final.py:44 (scale_factor=4) / loss.py:19 (size=H).  align_corners=False."""
    if size is not None:
        return F.interpolate(input=low, mode="bilinear", size=size)
    return F.interpolate(input=low, mode="bilinear", scale_factor=scale_factor)


def auxiliary_loss(low: Tensor, target: Tensor, ignore_index: int = -100) -> Tensor:
    """``AuxiliaryLoss.forward`` restated (loss.py:17-21): bilinear to label size, then
    ``nn.CrossEntropyLoss(ignore_index=..., reduction='mean')``."""
    B, H, W = target.shape
    up = F.interpolate(input=low, mode="bilinear", size=H)        # loss.py:19
    return F.cross_entropy(up, target, ignore_index=ignore_index)  # loss.py:20


def auxiliary_loss_and_grad(low: Tensor, target: Tensor, ignore_index: int = -100):
    """loss + d loss / d low via autograd (the reference's ``loss.backward()``,
    engine.py:100).  Returns (loss, grad_low, n_valid)."""
    low = low.detach().clone().requires_grad_(True)
    loss = auxiliary_loss(low, target, ignore_index)
    (g,) = torch.autograd.grad(loss, low)
    n_valid = int((target != ignore_index).sum())
    return loss.detach(), g, n_valid


def cosine_logits_backward(v: Tensor, t: Tensor, grad_logits: Tensor, normalize: bool = True,
                           logit_scale: float = 1.0):
    """d/dv, d/dt of <grad_logits, cosine_logits(v, t)> via autograd (engine.py:100)."""
    v = v.detach().clone().float().requires_grad_(True)
    t = t.detach().clone().float().requires_grad_(True)
    B, P, D = v.shape
    h, w = grad_logits.shape[-2:]
    out = cosine_logits(v, t, normalize=normalize, logit_scale=logit_scale, hw_shape=(h, w))
    gv, gt = torch.autograd.grad(out, (v, t), grad_outputs=grad_logits)
    return gv, gt


# --------------------------------------------------------------------------------------
# Stage 3: argmax + confusion matrix + mIoU                 (metrics.py:61-134)
# --------------------------------------------------------------------------------------
def contrastive_loss(outputs: Tensor, labels: Tensor, ignore_index: int = -100, num_classes: int = 151):
    """model/loss.py:45-64 restated term by term (no nn.CrossEntropyLoss call, so that the row-axis softmax of the
    'textual' term is explicit).  outputs [B, h*w, C], labels [B, h, w] -> (total, loss_visual, loss_textual).

    loss.py:50-51: out_textual = [B,h,w,C] view, out_visual = [B,C,h,w]; :54 one-hot float target [B,h,w,151];
    :58 CE(out_textual, one-hot): probabilities target, class axis = dim 1 (the image-row axis), 'mean' = divide by
    B * w * C; :59 CE(out_visual, labels): class-index target with ignore_index, mean over counted pixels."""
    B, hw, C = outputs.shape
    h = int(math.isqrt(hw))
    o = outputs.view(B, h, hw // h, C)
    onehot = F.one_hot(labels, num_classes=num_classes).to(outputs.dtype)        # raises on labels outside [0,151)
    logp_rows = o - torch.logsumexp(o, dim=1, keepdim=True)                         # softmax over y
    loss_textual = -(onehot * logp_rows).sum() / (B * (hw // h) * C)
    logp_cls = o - torch.logsumexp(o, dim=3, keepdim=True)                          # softmax over classes
    counted = labels != ignore_index
    picked = logp_cls.gather(3, labels.clamp(0, C - 1).unsqueeze(-1)).squeeze(-1)
    loss_visual = -(picked * counted).sum() / counted.sum()
    return (loss_textual + loss_visual) / 2, loss_visual, loss_textual


def upsample_tokens_bicubic4(dec_v: Tensor, h: int) -> Tensor:
    """model/model.py:42-44 restated (einops rearranges written as permutes): token-major features [B, h*w, C] ->
    bicubic x4 -> [B, 16*h*w, C]."""
    B, P, C = dec_v.shape
    w = P // h
    x = dec_v.permute(0, 2, 1).reshape(B, C, h, w)                       # "b (h w) c -> b c h w"
    x = F.interpolate(input=x, mode="bicubic", scale_factor=4)
    return x.reshape(B, C, 16 * P).permute(0, 2, 1).contiguous()         # "b c h w -> b (h w) c"


def argmax_reference(logits: Tensor) -> Tensor:
    """What ``JaccardIndex`` sees in the reference: ``argmax(Softmax2d(x), dim=class)``
    (metrics.py:92 feeds ``softmax2D(output)``; torchmetrics argmaxes float preds).
    logits: [..., C, H, W] -> [..., H, W] int64.  First index wins ties (torch.argmax)."""
    return torch.softmax(logits, dim=-3).argmax(dim=-3)


def argmax_logits(logits: Tensor) -> Tensor:
    """``argmax`` taken on the logits directly (softmax is strictly monotone, so this is
    the same decision except where fp32 softmax rounds a near-tie together; used by the
    tests to select tie-gap-safe inputs)."""
    return logits.argmax(dim=-3)


def confusion_matrix(pred: Tensor, target: Tensor, num_classes: int) -> Tensor:
    """torchmetrics ``_confusion_matrix_update`` restated: rows = target, cols = pred.
    int64 [C, C].  (parity unpinned - see module header.)"""
    p = pred.reshape(-1).to(torch.int64)
    t = target.reshape(-1).to(torch.int64)
    keep = (t >= 0) & (t < num_classes)
    idx = t[keep] * num_classes + p[keep]
    return torch.bincount(idx, minlength=num_classes * num_classes).reshape(num_classes, num_classes)


def jaccard_none(cm: Tensor) -> Tensor:
    """``JaccardIndex(num_classes, average='none')`` from a confusion matrix
    (metrics.py:64,85): per-class IoU, union==0 -> 0.0 (torchmetrics absent_score=0)."""
    cm = cm.to(torch.float32)
    inter = torch.diag(cm)
    union = cm.sum(0) + cm.sum(1) - inter
    iou = inter / union
    iou[union == 0] = 0.0
    return iou


def jaccard_macro(cm: Tensor, ignore_index: Optional[int]) -> Tensor:
    """``JaccardIndex(num_classes, ignore_index=i)`` default average='macro'
    (metrics.py:130): zero the ignore row, per-class IoU, drop the ignore class, mean
    (absent classes count as 0.0)."""
    cm = cm.clone()
    C = cm.shape[0]
    if ignore_index is not None and 0 <= ignore_index < C:
        cm[ignore_index] = 0
    iou = jaccard_none(cm)
    if ignore_index is not None and 0 <= ignore_index < C:
        iou = torch.cat([iou[:ignore_index], iou[ignore_index + 1:]])
    return iou.mean()


def per_image_miou_from_cm(cm: Tensor, label: Tensor, ignore_index: Optional[int]) -> Tensor:
    """metrics.py:91-97: IoU (average='none') indexed by the classes present in the label,
    minus ignore_index, mean."""
    iou = jaccard_none(cm)
    classes = label.unique()
    if ignore_index is None:
        return iou[classes.long()].mean(dim=0, keepdim=True)
    return iou[classes[classes != ignore_index].long()].mean(dim=0, keepdim=True)


def compute_mIOU(outputs: Tensor, labels: Tensor, n_cls: int, ignore_index: Optional[int] = 0,
                 argmax_fn=argmax_reference) -> dict:
    """metrics.py:82-102 restated.  Per image: bicubic x4 of the logits (:89), nearest x4
    of the labels (:90), JaccardIndex(average='none') on softmax (:92), mean over present
    classes != ignore_index (:94-97); then mean over images (:101)."""
    all_miou = []
    for i in range(len(labels)):
        output, label = outputs[i].unsqueeze(0), labels[i].unsqueeze(0)
        output = F.interpolate(input=output, mode="bicubic", scale_factor=4).squeeze(0)
        label = F.interpolate(input=label.view(-1, 1, label.shape[-1], label.shape[-1]).float(),
                              mode="nearest", scale_factor=4).squeeze().long()
        pred = argmax_fn(output)
        cm = confusion_matrix(pred, label, n_cls)
        all_miou.append(per_image_miou_from_cm(cm, label, ignore_index))
    return dict(mIOU_label=torch.concat(all_miou).mean().item())


def compute_gt_mIOU(outputs: Tensor, gt_list: List[Tensor], sizes, n_cls: int = 151,
                    ignore_index: Optional[int] = 0, argmax_fn=argmax_reference) -> dict:
    """metrics.py:61-79 restated: bicubic to each image's original size (:67)."""
    all_miou = []
    for i in range(len(gt_list)):
        size = tuple(int(s) for s in sizes[i])
        pred_l = F.interpolate(input=outputs[i].unsqueeze(0), mode="bicubic", size=size).squeeze(0)
        pred = argmax_fn(pred_l)
        cm = confusion_matrix(pred, gt_list[i], n_cls)
        all_miou.append(per_image_miou_from_cm(cm, gt_list[i], ignore_index))
    return dict(mIOU_gt=torch.concat(all_miou).mean().item())


def compute_mIOU_tensor(pred: Tensor, label: Tensor, n_cls: int, ignore_index: Optional[int] = 0,
                        argmax_fn=argmax_reference) -> float:
    """metrics.py:127-134 restated: dataset-level JaccardIndex(num_classes, ignore_index)."""
    p = argmax_fn(pred)
    cm = confusion_matrix(p, label, n_cls)
    return jaccard_macro(cm, ignore_index).item()


def segmentation_metrics(outputs, labels, gt_list, sizes, n_clas: int = 151, ignore_index: Optional[int] = 0) -> dict:
    """metrics.py:45-58."""
    m = {}
    m.update(compute_mIOU(outputs=outputs, labels=labels, n_cls=n_clas, ignore_index=ignore_index))
    m.update(compute_gt_mIOU(outputs=outputs, gt_list=gt_list, sizes=sizes, n_cls=n_clas, ignore_index=ignore_index))
    return m


def generate_masks(preds: Tensor, sizes) -> List[Tensor]:
    """utils.py:15-22: bicubic to original size then argmax(dim=0)."""
    up = [F.interpolate(input=p.unsqueeze(0), mode="bicubic", size=tuple(int(x) for x in s)).squeeze(0)
          for p, s in zip(preds, sizes)]
    return [x.argmax(dim=0) for x in up]


def pixel_accuracy(cm: Tensor, ignore_index: Optional[int]) -> float:
    """Not in the reference: trace/sum over non-ignored target rows."""
    cm = cm.clone()
    if ignore_index is not None and 0 <= ignore_index < cm.shape[0]:
        cm[ignore_index] = 0
    tot = cm.sum().item()
    return float(torch.diag(cm).sum().item()) / tot if tot else 0.0


# --------------------------------------------------------------------------------------
# Whole hot path on CPU (the bench's cpu_baseline / --impl reference leg)
# --------------------------------------------------------------------------------------
def head_step(v: Tensor, t: Tensor, labels: Tensor, ignore_index: int, n_cls: int,
              normalize: bool = True, logit_scale: float = 1.0, backward: bool = True,
              hw_shape=None):
    """One pass of the reference hot path on CPU, fp32:
    final.py:41-43 -> loss.py:17-21 (AuxiliaryLoss: bilinear to label size + CE) ->
    engine.py:100 backward -> metrics.py:127-134-style argmax/confusion matrix on the
    bilinear-upsampled score map (final.py:44 'outputs')."""
    v = v.detach().float().requires_grad_(backward)
    t = t.detach().float().requires_grad_(backward)
    low = cosine_logits(v, t, normalize=normalize, logit_scale=logit_scale, hw_shape=hw_shape)
    B, H, W = labels.shape
    up = F.interpolate(input=low, mode="bilinear", size=H)
    loss = F.cross_entropy(up, labels, ignore_index=ignore_index)
    gv = gt = None
    if backward:
        gv, gt = torch.autograd.grad(loss, (v, t))
    with torch.no_grad():
        pred = argmax_reference(up)
        cm = confusion_matrix(pred, labels, n_cls)
        miou = jaccard_macro(cm, ignore_index)
    return dict(loss=loss.detach(), grad_v=gv, grad_t=gt, confmat=cm, miou=miou, low=low.detach())
