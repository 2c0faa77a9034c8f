"""Mirror of the reference ``model/decoder.py`` call surface (decoder.py:9-33).

These are the thin ``nn.TransformerDecoder*`` subclasses that sit UPSTREAM of the head; they are
attention stacks, not one of the four hot-path stages, so they stay plain
PyTorch pass-throughs with the reference's constructor signatures.  ``FTNDecoder`` /
``SRTransformerDecoder`` (decoder.py:36-134) override ``_sa_block`` without the ``is_causal``
argument torch >= 2.0 passes, i.e. they do not run in the reference either on this torch; they are
out of scope and deliberately absent.
"""
from typing import Callable, Optional, Union

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
