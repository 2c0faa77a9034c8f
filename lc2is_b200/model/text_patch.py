"""Mirror of the reference ``model/text_patch.py:4-18``.

Same constructor, same submodule names (``textual``, ``visual``) so reference state_dicts load
(``pixel_patch.textual.*`` / ``pixel_patch.visual.*``), same return ORDER: text first.
The two projections are plain library GEMMs (cuBLAS through ``nn.Linear``); they feed the head
kernels and are not part of the four hand-written stages (SURVEY 8f-1 marks fusing
``visual`` in front of the logits GEMM as the next widening step).
"""
from torch import Tensor, nn


class TextToPatch(nn.Module):

    def __init__(self, img_in: int, text_in: int, out: int = 512) -> None:
        super().__init__()
        # img   (batch, patches, img_in)  --> (batch, patches, out)
        # text  (classes, text_in)        --> (classes, out)
        self.textual = nn.Linear(in_features=text_in, out_features=out)
        self.visual = nn.Linear(in_features=img_in, out_features=out)

    def forward(self, img: Tensor, text: Tensor) -> tuple[Tensor, Tensor]:
        return self.textual(text), self.visual(img)
