"""Mirror of the reference ``model/loss.py`` (same class names, ctor signatures, return values).

``AuxiliaryLoss`` (loss.py:12-21) runs on the fused K2 kernel: bilinear upsample to the label
size + softmax cross-entropy + their backward in one pass, with no upsampled tensor in memory.
It is also the drop-in for ``nn.CrossEntropyLoss`` applied to ``F.interpolate(score, 'bilinear')``
(final.py:44 + engine.py:94): feed it the low-resolution score map instead.
``ContrastiveLoss`` / ``NPairLoss`` (loss.py:23-64) are kept as plain PyTorch (next).
"""
from typing import Callable, Optional, Union

import numpy as np
import torch
import torch.nn.functional as F
from einops import rearrange
from torch import Tensor, nn

from .. import ops


class _UpsampleCE(torch.autograd.Function):

    @staticmethod
    def forward(ctx, low: Tensor, target: Tensor, ignore_index: int, reduction: str):
        need_grad = ctx.needs_input_grad[0]
        low32 = low.float()
        n_valid = ops.count_valid(target, ignore_index)
        gscale = ops.mean_scale(n_valid) if reduction == "mean" else None
        loss_sum, grad, _ = ops.upsample_ce(low32, target, ignore_index, gscale, want_grad=need_grad)
        loss = ops.finalize_loss(loss_sum, n_valid) if reduction == "mean" else loss_sum.float()
        if need_grad:
            ctx.save_for_backward(grad)
        ctx.in_dtype = low.dtype
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        (grad,) = ctx.saved_tensors
        return (grad * grad_out).to(ctx.in_dtype), None, None, None


class AuxiliaryLoss(nn.CrossEntropyLoss):
    """loss.py:12-21.  forward(input [B,C,h,w], target [B,H,W] int64) -> scalar."""

    def __init__(self, weight: Optional[Tensor] = None, size_average=None, ignore_index: int = -100, reduce=None,
                 reduction: str = 'mean', label_smoothing: float = 0) -> None:
        super().__init__(weight, size_average, ignore_index, reduce, reduction, label_smoothing)
        if weight is not None or label_smoothing != 0 or self.reduction not in ("mean", "sum"):
            raise NotImplementedError(
                "the B200 AuxiliaryLoss kernel covers the reference's usage (no class weights, no label "
                "smoothing, reduction 'mean'|'sum'); there is deliberately no silent PyTorch fallback")

    def forward(self, input: Tensor, target: Tensor) -> Tensor:
        B, H, W = target.shape
        if W != H:
            # the reference passes size=H (an int): the output is H x H regardless of W (loss.py:19)
            raise ValueError("AuxiliaryLoss follows loss.py:19 (size=H): labels must be square")
        return _UpsampleCE.apply(input, target, self.ignore_index, self.reduction)


class NPairLoss(nn.Module):
    """loss.py:23-37 (PyTorch pass-through)."""

    def __init__(self, reduction: Union[Callable, None] = torch.mean) -> None:
        super().__init__()
        self.reduction = reduction

    def forward(self, x: Tensor, x_pos: Tensor, x_neg: Tensor):
        pos = torch.matmul(x, x_pos.transpose(0, 1))
        neg = torch.matmul(x, x_neg.transpose(0, 1)).sum(-1, keepdim=True)
        res = (pos / (pos + neg)).sum(-1)
        if self.reduction:
            res = self.reduction(res)
        return res


class ContrastiveLoss(nn.Module):
    """loss.py:39-64 (PyTorch pass-through, including the 151-way one-hot whose class axis is the
    image-row axis, loss.py:51-60)."""

    def __init__(self, weight: Optional[Tensor] = None, size_average=None, ignore_index: int = -100, reduce=None,
                 reduction: str = 'mean', label_smoothing: float = 0) -> None:
        super().__init__()
        self.criterion = nn.CrossEntropyLoss(weight, size_average, ignore_index, reduce, reduction, label_smoothing)

    def forward(self, outputs: Tensor, labels: Tensor):
        H = int(np.sqrt(outputs.shape[1]).item())
        out_textual = rearrange(outputs, "b (h w) c -> b h w c", h=H)
        out_visual = rearrange(outputs.transpose(-2, -1), "b c (h w) -> b c h w", h=H)
        label_textual = F.one_hot(labels, num_classes=151).float()
        label_visual = labels
        loss_textual = self.criterion(input=out_textual, target=label_textual)
        loss_visual = self.criterion(input=out_visual, target=label_visual)
        return (loss_textual + loss_visual) / 2, loss_visual, loss_textual
