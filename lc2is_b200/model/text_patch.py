"""Mirror of the reference ``model/text_patch.py:4-18``.

Same constructor, same submodule names (``textual``, ``visual``) so reference state_dicts load
(``pixel_patch.textual.*`` / ``pixel_patch.visual.*``), same return ORDER: text first.

The visual projection ([B*P, img_in] x [img_in, out]: 5x the FLOPs of the logits GEMM and genuinely tensor-bound,
SURVEY 8f-1) runs its FORWARD on the logits GEMM's tcgen05 / TMEM / TMA pipeline (``lc2is_linear_fwd``: bf16 operands,
fp32 accumulate, fused bias) when the input is a CUDA tensor; its backward (dX, dW, db) and the tiny text projection
([C, text_in]) are library GEMMs through torch.  CPU tensors go through ``nn.Linear`` unchanged - the projections sit
upstream of the head hot path (the head kernels themselves have no CPU path).
"""
import torch
from torch import Tensor, nn

from .. import ops


class _LinearTC(torch.autograd.Function):
    """y = x W^T + b, forward on lc2is_linear_fwd (bf16 operands), backward with torch matmuls on the same operands."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Tensor):
        K = x.shape[-1]
        xb = x.reshape(-1, K).to(torch.bfloat16).contiguous()
        wb = weight.to(torch.bfloat16).contiguous()
        out_dtype = torch.bfloat16 if x.dtype == torch.bfloat16 else torch.float32
        y = ops.linear_fwd(xb, wb, None if bias is None else bias.float().contiguous(), out_dtype)
        ctx.save_for_backward(xb, wb)
        ctx.meta = (x.shape, x.dtype, weight.dtype, bias is not None)
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, gy: Tensor):
        xb, wb = ctx.saved_tensors
        shape, xdt, wdt, has_bias = ctx.meta
        g = gy.reshape(-1, gy.shape[-1]).to(torch.bfloat16)
        gx = (g @ wb).view(shape).to(xdt) if ctx.needs_input_grad[0] else None
        gw = (g.t() @ xb).to(wdt) if ctx.needs_input_grad[1] else None
        gb = gy.reshape(-1, gy.shape[-1]).float().sum(0) if (has_bias and ctx.needs_input_grad[2]) else None
        return gx, gw, gb


class TextToPatch(nn.Module):

    def __init__(self, img_in: int, text_in: int, out: int = 512) -> None:
        super().__init__()
        # img   (batch, patches, img_in)  --> (batch, patches, out)
        # text  (classes, text_in)        --> (classes, out)
        self.textual = nn.Linear(in_features=text_in, out_features=out)
        self.visual = nn.Linear(in_features=img_in, out_features=out)

    def forward(self, img: Tensor, text: Tensor) -> tuple[Tensor, Tensor]:
        t_feature = self.textual(text)
        lin = self.visual
        if img.is_cuda and lin.in_features % 64 == 0 and lin.out_features % 16 == 0:
            v_feature = _LinearTC.apply(img, lin.weight, lin.bias)
        else:
            v_feature = lin(img)
        return t_feature, v_feature
