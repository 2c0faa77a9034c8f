"""Seeded synthetic ADE20K-shaped head inputs.  CPU tensors; callers move them.

The reference needs the ADE20K download and HF weights to make real inputs
(data/ade20k/prepare_ade20k.py, model/encoder.py); neither exists offline, so the
bench and the tests drive the head with tensors of the same shape, dtype and label
contract (int64 [B,H,W], values 0..C-1, 0 = "none": data/collator.py:91,
data/dataset.py:46-49).
"""
from __future__ import annotations

import os

import torch

SEED = 1024  # the reference's default seed, evaluate.py:24

_PROTO_PATH = os.path.join(os.path.dirname(__file__), "data", "ade20k_prototypes.pt")
PROTO_SHA256_RAW = "7d29cfae8b153fe0bab10b15566a7fcfd0b8daa4dd95faff8779ad2ed5075a36"


def load_prototypes() -> torch.Tensor:
    """The reference's bundled [151,512] fp32 class prototypes (model/model.py:22)."""
    return torch.load(_PROTO_PATH).detach().clone().float()


def make_prototypes(C: int, D: int = 512, seed: int = SEED) -> torch.Tensor:
    """[C,D] text embeddings: the bundled file for C=151, its rows 1.. for C=150
    (drops "none"), otherwise randn*1.05 (matches the file's std)."""
    if D == 512 and C == 151:
        return load_prototypes()
    if D == 512 and C == 150:
        return load_prototypes()[1:].contiguous()
    g = torch.Generator().manual_seed(seed + 7)
    return torch.randn(C, D, generator=g) * 1.05


def make_patch_embeddings(B: int, hw: int, D: int = 512, seed: int = SEED, dtype=torch.bfloat16) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, hw, D, generator=g).to(dtype)


def make_labels(B: int, H: int, W: int, C: int, seed: int = SEED, block: int = 16,
                noise: float = 0.05, ignore_frac: float = 0.0, ignore_index: int = 0) -> torch.Tensor:
    """Blocky ADE-like label maps: randint on a (H/block, W/block) grid, nearest x block,
    then `noise` of the pixels re-drawn uniformly; optionally force `ignore_frac` zeros."""
    g = torch.Generator().manual_seed(seed + 1)
    gh, gw = max(H // block, 1), max(W // block, 1)
    coarse = torch.randint(0, C, (B, gh, gw), generator=g)
    lab = coarse.repeat_interleave(block, 1).repeat_interleave(block, 2)[:, :H, :W]
    if lab.shape[1] != H or lab.shape[2] != W:
        pad = torch.randint(0, C, (B, H, W), generator=g)
        pad[:, :lab.shape[1], :lab.shape[2]] = lab
        lab = pad
    lab = lab.contiguous().clone()
    if noise > 0:
        m = torch.rand(B, H, W, generator=g) < noise
        r = torch.randint(0, C, (B, H, W), generator=g)
        lab[m] = r[m]
    if ignore_frac > 0:
        m = torch.rand(B, H, W, generator=g) < ignore_frac
        lab[m] = ignore_index
    return lab.to(torch.int64)


def make_dyadic_logits(B: int, C: int, h: int, w: int, seed: int = SEED) -> torch.Tensor:
    """Exactness set: randint(-1024,1025)/256 fp32 - bilinear x4/x16 of these is exact
    in fp32 in any evaluation order, so fused-upsample argmax is order independent."""
    g = torch.Generator().manual_seed(seed + 3)
    return torch.randint(-1024, 1025, (B, C, h, w), generator=g).float() / 256.0
