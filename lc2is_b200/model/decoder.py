"""Mirror of the reference ``model/decoder.py`` call surface (decoder.py:9-33).

These are the thin ``nn.TransformerDecoder*`` subclasses that sit UPSTREAM of the head; they are
attention stacks, not one of the four hot-path stages, so they stay plain
PyTorch pass-throughs with the reference's constructor signatures and parameter names (state_dicts load).
``SRTransformerDecoder`` (decoder.py:113-134) overrides ``_sa_block`` without the ``is_causal`` argument torch >= 2.0
passes, so it - and ``FTNBlock`` / ``FTNDecoder`` built on it - raise ``TypeError`` in the reference on this torch; the
mirrors accept (and ignore) that keyword, everything else follows the reference's data flow.
(The reference hands ``device, dtype`` to ``nn.TransformerDecoderLayer.__init__`` positionally; since torch 2.1 that slot
is ``bias``, so on this torch the reference's layers silently lose their biases.  The mirrors pass them by keyword, i.e.
they keep the layout of the torch 1.x the reference was written for.)
"""
from typing import Callable, List, Optional, Union

import torch
import torch.nn.functional as F
from torch import Tensor, nn


class DecoderLayer(nn.TransformerDecoderLayer):
    """decoder.py:9-13: cross-attention keys/values of width ``d_kv``."""

    def __init__(self, d_model: int, d_kv: int, nhead: int, dim_feedforward: int = 2048, dropout: float = 0,
                 activation: Union[str, Callable[[Tensor], Tensor]] = F.relu, layer_norm_eps: float = 0.00001,
                 batch_first: bool = False, norm_first: bool = False, device=None, dtype=None) -> None:
        super().__init__(d_model, nhead, dim_feedforward, dropout, activation, layer_norm_eps, batch_first,
                         norm_first, device=device, dtype=dtype)
        self.multihead_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout, batch_first=batch_first,
                                                    kdim=d_kv, vdim=d_kv, device=device, dtype=dtype)


class DecoderBlock(nn.TransformerDecoder):
    """decoder.py:15-21."""

    def __init__(self, decoder_layer, num_layers, norm=None):
        super().__init__(decoder_layer, num_layers, norm)

    def forward(self, tgt: Tensor, memory: Tensor, tgt_mask: Optional[Tensor] = None,
                memory_mask: Optional[Tensor] = None, tgt_key_padding_mask: Optional[Tensor] = None,
                memory_key_padding_mask: Optional[Tensor] = None) -> Tensor:
        return super().forward(tgt, memory, tgt_mask, memory_mask, tgt_key_padding_mask, memory_key_padding_mask)


class PromptLayer(nn.TransformerDecoderLayer):
    """decoder.py:24-28."""

    def __init__(self, d_model: int, d_kv: int, nhead: int, dim_feedforward: int = 2048, dropout: float = 0.1,
                 activation: Union[str, Callable[[Tensor], Tensor]] = F.relu, layer_norm_eps: float = 0.00001,
                 batch_first: bool = False, norm_first: bool = False, device=None, dtype=None) -> None:
        super().__init__(d_model, nhead, dim_feedforward, dropout, activation, layer_norm_eps, batch_first,
                         norm_first, device=device, dtype=dtype)
        self.multihead_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout, batch_first=batch_first,
                                                    kdim=d_kv, vdim=d_kv, device=device, dtype=dtype)


class PromptDecoder(nn.TransformerDecoder):
    """decoder.py:30-33."""

    def __init__(self, decoder_layer, num_layers, norm=None):
        super().__init__(decoder_layer, num_layers, norm)


def _tokens_to_map(x: Tensor) -> Tensor:
    """[B, h*h, C] -> [B, C, h, h] (square token grids, as the reference assumes)."""
    b, n, c = x.shape
    side = int(n ** 0.5)
    return x.transpose(1, 2).reshape(b, c, side, side)


def _map_to_tokens(x: Tensor) -> Tensor:
    return x.flatten(2).transpose(1, 2)


def _upsample_tokens(x: Tensor, factor: int) -> Tensor:
    return _map_to_tokens(F.interpolate(_tokens_to_map(x), mode="bilinear", scale_factor=factor))


class SRTransformerDecoder(nn.TransformerDecoderLayer):
    """decoder.py:113-134: decoder layer whose SELF-attention keys / values are the token map shrunk by a strided
    convolution (``sr``) and layer-normed (``norm``) - spatial-reduction attention."""

    def __init__(self, d_model: int, nhead: int, sr_ratio: int = 1, dim_feedforward: int = 2048, dropout: float = 0.1,
                 activation: Union[str, Callable[[Tensor], Tensor]] = F.relu, layer_norm_eps: float = 0.00001,
                 batch_first: bool = False, norm_first: bool = False, device=None, dtype=None) -> None:
        super().__init__(d_model, nhead, dim_feedforward, dropout, activation, layer_norm_eps, batch_first,
                         norm_first, device=device, dtype=dtype)
        self.sr_ratio = sr_ratio
        self.sr = nn.Conv2d(d_model, d_model, kernel_size=sr_ratio, stride=sr_ratio)
        self.norm = nn.LayerNorm(d_model)

    def _sa_block(self, x: Tensor, attn_mask: Optional[Tensor], key_padding_mask: Optional[Tensor],
                  is_causal: bool = False) -> Tensor:
        kv = self.norm(_map_to_tokens(self.sr(_tokens_to_map(x)))) if self.sr_ratio > 1 else x
        out = self.self_attn(x, kv, kv, attn_mask=attn_mask, key_padding_mask=key_padding_mask, need_weights=False)[0]
        return self.dropout1(out)


class FTNBlock(nn.Module):
    """decoder.py:96-111: attention block, then bilinear x``upsample`` of the token map."""

    def __init__(self, attention_block: nn.Module, upsample: int = 2) -> None:
        super().__init__()
        self.attention_block = attention_block
        self.upsample = upsample

    def forward(self, tgt: Tensor, memory: Tensor) -> Tensor:
        return _upsample_tokens(self.attention_block(tgt=tgt, memory=memory), self.upsample)


class FTNDecoder(nn.Module):
    """decoder.py:36-94: top-down pyramid over four encoder stages; stages 2-4 pass through 1 / 2 / 3 upsampling
    attention blocks against the text tokens and the four maps (all at stage-1 resolution by then) are summed."""

    def __init__(self, in_dims: List[int], dim: int, dropout: float = 0.1) -> None:
        super().__init__()
        self.linear_stage_2 = nn.Linear(in_dims[2], in_dims[1])
        self.linear_stage_3 = nn.Linear(in_dims[3], in_dims[2])
        for i in range(4):
            setattr(self, f"linear2_stage_{i + 1}", nn.Linear(in_dims[i], dim))

        def blocks(n):
            return nn.ModuleList(FTNBlock(SRTransformerDecoder(d_model=dim, nhead=8, sr_ratio=2, dropout=dropout,
                                                               batch_first=True)) for _ in range(n))
        self.attention_stage_2, self.attention_stage_3, self.attention_stage_4 = blocks(1), blocks(2), blocks(3)

    def forward(self, visual: List[Tensor], textual: Tensor) -> Tensor:
        s4 = visual[3]
        s3 = self.linear_stage_3(_upsample_tokens(s4, 2))
        s2 = self.linear_stage_2(_upsample_tokens(s3, 2))
        maps = [self.linear2_stage_1(visual[0]), self.linear2_stage_2(s2), self.linear2_stage_3(s3),
                self.linear2_stage_4(s4)]
        for i, stack in ((1, self.attention_stage_2), (2, self.attention_stage_3), (3, self.attention_stage_4)):
            for block in stack:
                maps[i] = block(tgt=maps[i], memory=textual)
        return torch.stack(maps, dim=0).sum(dim=0)
