"""Mirror of the reference ``model/text_patch.py:4-18``.

Same constructor, same submodule names (``textual``, ``visual``) so reference state_dicts load
(``pixel_patch.textual.*`` / ``pixel_patch.visual.*``), same return ORDER: text first.

The visual projection ([B*P, img_in] x [img_in, out]: 5x the FLOPs of the logits GEMM and genuinely tensor-bound,
SURVEY 8f-1) runs on the logits GEMM's tcgen05 / TMEM / TMA pipeline in both directions when its input is a CUDA bf16
tensor (what the encoders hand over under autocast) - forward ``lc2is_linear_fwd`` (bf16 operands, fp32 accumulate, fused
bias), backward ``lc2is_linear_bwd`` (dX through the same pipeline on the transposed weight; dW as a split-K GEMM with
fp32 accumulation and fp32 output, so fp32 master weights receive fp32 gradients; db a column-sum stream).  fp32 inputs
keep ``nn.Linear``'s fp32 arithmetic unless ``tensor_cores=True`` asks for the bf16-operand path explicitly
(tolerance: operands rounded to bf16, 2^-9 relative per element).  The tiny text projection ([C, text_in]) and CPU
tensors go through ``nn.Linear`` unchanged - the projections sit upstream of the head hot path.
"""
import torch
from torch import Tensor, nn

from .. import ops


class _LinearTC(torch.autograd.Function):
    """y = x W^T + b: forward on lc2is_linear_fwd, backward on lc2is_linear_bwd (bf16 operands, fp32 accumulation)."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Tensor):
        K = x.shape[-1]
        xb = x.reshape(-1, K).to(torch.bfloat16).contiguous()
        wb = weight.to(torch.bfloat16).contiguous()
        out_dtype = torch.bfloat16 if x.dtype == torch.bfloat16 else torch.float32
        y = ops.linear_fwd(xb, wb, None if bias is None else bias.float().contiguous(), out_dtype)
        ctx.save_for_backward(xb, wb)
        ctx.meta = (x.shape, x.dtype, weight.dtype, bias is not None, None if bias is None else bias.dtype)
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, gy: Tensor):
        xb, wb = ctx.saved_tensors
        shape, xdt, wdt, has_bias, bdt = ctx.meta
        g = gy.reshape(-1, gy.shape[-1]).to(torch.bfloat16).contiguous()
        need = ctx.needs_input_grad
        gx, gw, gb = ops.linear_bwd(g, xb, wb, need_gx=need[0], need_gw=need[1], need_gb=has_bias and need[2],
                                    gx_dtype=torch.bfloat16 if xdt == torch.bfloat16 else torch.float32)
        return (gx.view(shape).to(xdt) if gx is not None else None, gw.to(wdt) if gw is not None else None,
                gb.to(bdt) if gb is not None else None)


class TextToPatch(nn.Module):

    def __init__(self, img_in: int, text_in: int, out: int = 512, tensor_cores: bool = False) -> None:
        super().__init__()
        self.tensor_cores = tensor_cores      # bf16-operand tcgen05 path for fp32 inputs too (bf16 inputs always take it)
        # img   (batch, patches, img_in)  --> (batch, patches, out)
        # text  (classes, text_in)        --> (classes, out)
        self.textual = nn.Linear(in_features=text_in, out_features=out)
        self.visual = nn.Linear(in_features=img_in, out_features=out)

    def forward(self, img: Tensor, text: Tensor) -> tuple[Tensor, Tensor]:
        t_feature = self.textual(text)
        lin = self.visual
        if img.is_cuda and (img.dtype == torch.bfloat16 or self.tensor_cores) \
                and lin.in_features % 64 == 0 and lin.out_features % 64 == 0:
            v_feature = _LinearTC.apply(img, lin.weight, lin.bias)
        else:
            v_feature = lin(img)
        return t_feature, v_feature
