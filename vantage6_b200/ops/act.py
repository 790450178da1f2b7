"""Fused activation ops (csrc/act.cu): SwiGLU ``silu(g) * u`` forward / backward in one pass each."""
from __future__ import annotations

import torch

from . import count, native, stream_ptr


def _fast(g: torch.Tensor, u: torch.Tensor) -> bool:
    return (g.is_cuda and g.dtype == torch.bfloat16 and u.dtype == torch.bfloat16 and g.shape == u.shape
            and g.is_contiguous() and u.is_contiguous() and g.numel() % 8 == 0 and g.numel() >= 8
            and g.data_ptr() % 16 == 0 and u.data_ptr() % 16 == 0)


class _SwiGLUFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g, u):
        h = torch.empty_like(g)
        count(1)
        native().swiglu_fwd(g.data_ptr(), u.data_ptr(), h.data_ptr(), g.numel(), stream_ptr())
        ctx.save_for_backward(g, u)
        return h

    @staticmethod
    def backward(ctx, dh):
        g, u = ctx.saved_tensors
        if not dh.is_contiguous():
            dh = dh.contiguous()
        dg, du = torch.empty_like(g), torch.empty_like(u)
        count(1)
        native().swiglu_bwd(dh.data_ptr(), g.data_ptr(), u.data_ptr(), dg.data_ptr(), du.data_ptr(), g.numel(), stream_ptr())
        return dg, du


def swiglu(g: torch.Tensor, u: torch.Tensor) -> torch.Tensor:
    """``silu(g) * u`` (Llama MLP gate); the hand-written kernels on dense CUDA bf16 tensors, PyTorch ops elsewhere."""
    if _fast(g, u):
        return _SwiGLUFn.apply(g, u)
    return torch.nn.functional.silu(g) * u
