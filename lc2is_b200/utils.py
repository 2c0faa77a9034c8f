"""Mirror of the hot-path-adjacent helper in the reference ``utils.py``."""
from typing import List

import torch

from . import ops


def generate_masks(preds: torch.Tensor, sizes: torch.Tensor) -> List[torch.Tensor]:
    """utils.py:15-22: bicubic to each image's original size, then argmax(dim=0) -> int64 [H_i,W_i].
    Runs K3's fused resize+argmax (the confusion matrix it also produces is discarded)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    masks = []
    for pred, size in zip(preds, sizes):
        H, W = (int(x) for x in size)
        o = pred.unsqueeze(0).to(dev, torch.float32).contiguous()
        dummy = torch.zeros(1, 1, 1, dtype=torch.int64, device=dev)
        _, _, m = ops.argmax_confmat(o, dummy, want_pred=True, size=(H, W), mode="bicubic")
        masks.append(m[0])
    return masks


def count_params(model: torch.nn.Module, trainable: bool = False):
    """utils.py:6-13."""
    if trainable:
        return sum(p.numel() for p in model.parameters() if p.requires_grad) / 1e6
    return sum(p.numel() for p in model.parameters()) / 1e6
