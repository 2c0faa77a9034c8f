"""Mirror of the hot-path-adjacent helper in the reference ``utils.py``."""
from typing import List

import torch

from . import ops


def generate_masks(preds: torch.Tensor, sizes: torch.Tensor) -> List[torch.Tensor]:
    """utils.py:15-22: bicubic to each image's original size, then argmax(dim=0) -> int64 [H_i,W_i].
    One launch of the ragged resize+argmax kernel for the whole batch (lc2is_argmax_confmat_ragged)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    if isinstance(preds, (list, tuple)):
        preds = torch.stack(list(preds))
    sz = [(int(s[0]), int(s[1])) for s in sizes]
    low = preds.to(dev, torch.float32).contiguous()
    _, _, flat, desc = ops.argmax_confmat_ragged(low, sz, None, mode="bicubic", want_pred=True)
    return [flat[int(o):int(o) + H * W].view(H, W) for (o, H, W, _), (H, W) in zip(desc.tolist(), sz)]


def count_params(model: torch.nn.Module, trainable: bool = False):
    """utils.py:6-13."""
    if trainable:
        return sum(p.numel() for p in model.parameters() if p.requires_grad) / 1e6
    return sum(p.numel() for p in model.parameters()) / 1e6
