"""ResNet stem helpers on channels-last bf16 activations (csrc/pool.cu): 3x3/stride-2/pad-1 max pooling
(forward saves the arg-max position, backward is a deterministic gather) and the fused
uint8-NCHW -> normalised-bf16-NHWC image transform.  Anywhere the kernels do not apply (CPU, other
dtypes / layouts / pooling shapes) the stock PyTorch ops run."""
from __future__ import annotations

from typing import Sequence

import torch
from torch import nn

from . import count, native, stream_ptr


def _fast(x: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4 and x.shape[1] % 8 == 0
            and x.is_contiguous(memory_format=torch.channels_last))


class _MaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        N, C, H, W = x.shape
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        y = torch.empty((N, C, Ho, Wo), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
        idx = torch.empty((N, Ho, Wo, C), device=x.device, dtype=torch.uint8)
        count(1)
        native().maxpool3x3s2_fwd(x.data_ptr(), y.data_ptr(), idx.data_ptr(), N, H, W, C, stream_ptr())
        ctx.save_for_backward(idx)
        ctx.shape = (N, C, H, W)
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        N, C, H, W = ctx.shape
        if not dy.is_contiguous(memory_format=torch.channels_last):
            dy = dy.contiguous(memory_format=torch.channels_last)
        dx = torch.empty((N, C, H, W), device=dy.device, dtype=dy.dtype, memory_format=torch.channels_last)
        count(1)
        native().maxpool3x3s2_bwd(dy.data_ptr(), idx.data_ptr(), dx.data_ptr(), N, H, W, C, stream_ptr())
        return dx


class MaxPool3x3s2(nn.MaxPool2d):
    """``nn.MaxPool2d(3, stride=2, padding=1)`` with the hand-written NHWC kernels on the fast path."""

    def __init__(self):
        super().__init__(3, stride=2, padding=1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if _fast(x):
            return _MaxPoolFn.apply(x)
        return super().forward(x)


def image_normalize(img: torch.Tensor, mean: Sequence[float], std: Sequence[float]) -> torch.Tensor:
    """uint8 [N,3,H,W] -> (img - mean) / std as a channels-last bf16 [N,3,H,W] tensor (one pass)."""
    N, C, H, W = img.shape
    assert C == 3
    if img.is_cuda and img.dtype == torch.uint8 and img.is_contiguous() and (H * W) % 4 == 0:
        out = torch.empty((N, 3, H, W), device=img.device, dtype=torch.bfloat16, memory_format=torch.channels_last)
        count(1)
        native().image_normalize(img.data_ptr(), out.data_ptr(), N, H * W, float(mean[0]), float(mean[1]), float(mean[2]),
                                 1.0 / float(std[0]), 1.0 / float(std[1]), 1.0 / float(std[2]), stream_ptr())
        return out
    m = torch.tensor(list(mean), device=img.device, dtype=torch.float32).view(1, 3, 1, 1)
    s = torch.tensor(list(std), device=img.device, dtype=torch.float32).view(1, 3, 1, 1)
    return ((img.float() - m) / s).contiguous(memory_format=torch.channels_last)
