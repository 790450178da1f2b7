"""ResNet stem helpers on channels-last bf16 activations (csrc/pool.cu): 3x3/stride-2/pad-1 max pooling
(forward saves the arg-max position, backward is a deterministic gather) and the fused
uint8-NCHW -> normalised-bf16-NHWC image transform.  Anywhere the kernels do not apply (CPU, other
dtypes / layouts / pooling shapes) the stock PyTorch ops run."""
from __future__ import annotations

from typing import Optional, Sequence

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


def _stem_tc() -> bool:
    import os

    return os.environ.get("V6B200_CONV", "tc") == "tc" and os.environ.get("V6B200_STEM", "tc") == "tc"


class _StemS2DFn(torch.autograd.Function):
    """7x7/s2/p3 stem convolution on 3 channels as a 4x4/s1 convolution on the 16-channel space-to-depth image
    (csrc/pool.cu).  ``weight`` is the fp32 master filter ([O,3,7,7], channels-last memory), ``w_src`` the tensor the
    filter values are read from (the bf16 shadow view when attached)."""

    @staticmethod
    def forward(ctx, xs, weight, w_src, bn=None):
        O = weight.shape[0]
        ws = torch.empty((O, 16, 4, 4), device=xs.device, dtype=torch.bfloat16, memory_format=torch.channels_last)
        src = weight.detach() if w_src is None else w_src
        count(1)
        native().stem_weight_s2d(src.data_ptr(), src.dtype == torch.bfloat16, ws.data_ptr(), O, stream_ptr())
        ctx.tc = _stem_tc() and O % 64 == 0
        if ctx.tc:      # tcgen05 implicit GEMM through an overlapping-window im2col map (ops/conv.py::stem_fprop)
            from . import conv as C

            y = C.stem_fprop(xs, ws, bn=bn)       # bn: BatchNorm statistics of the output from the same launch
        else:
            y = torch.ops.aten.convolution(xs, ws, None, (1, 1), (0, 0), (1, 1), False, (0, 0), 1)
        ctx.save_for_backward(xs, ws)
        ctx.weight = weight
        return y

    @staticmethod
    def backward(ctx, dy):
        xs, ws = ctx.saved_tensors
        w = ctx.weight
        g = w.grad
        direct = g is not None and g.dtype == torch.float32 and g.is_contiguous(memory_format=torch.channels_last)
        dw = g if direct else torch.empty(w.shape, device=w.device, dtype=torch.float32, memory_format=torch.channels_last)
        if ctx.tc:
            from . import conv as C

            if not dy.is_contiguous(memory_format=torch.channels_last):
                dy = dy.contiguous(memory_format=torch.channels_last)
            dws32 = torch.zeros((w.shape[0], 4, 4, 16), device=w.device, dtype=torch.float32)
            C.stem_wgrad(dy, xs, dws32)
            count(1)
            native().stem_wgrad_d2s_f32(dws32.data_ptr(), dw.data_ptr(), w.shape[0], direct, stream_ptr())
            return (None, None, None, None) if direct else (None, dw, None, None)
        _, dws, _ = torch.ops.aten.convolution_backward(dy, xs, ws, None, (1, 1), (0, 0), (1, 1), False, (0, 0), 1,
                                                        (False, True, False))
        if not dws.is_contiguous(memory_format=torch.channels_last):
            dws = dws.contiguous(memory_format=torch.channels_last)
        count(1)
        native().stem_wgrad_d2s(dws.data_ptr(), dw.data_ptr(), w.shape[0], direct, stream_ptr())     # direct: straight into the flat grads
        return (None, None, None, None) if direct else (None, dw, None, None)


def stem_s2d_supported(img: torch.Tensor, conv: nn.Conv2d) -> bool:
    return (img.is_cuda and img.dtype == torch.uint8 and img.dim() == 4 and img.shape[1] == 3 and img.is_contiguous()
            and img.shape[2] % 2 == 0 and img.shape[3] % 2 == 0 and tuple(conv.kernel_size) == (7, 7)
            and tuple(conv.stride) == (2, 2) and tuple(conv.padding) == (3, 3) and conv.in_channels == 3
            and conv.bias is None and conv.weight.dtype == torch.float32
            and conv.weight.is_contiguous(memory_format=torch.channels_last))


def stem_s2d(img: torch.Tensor, conv: nn.Conv2d, mean: Sequence[float], std: Sequence[float], bn: Optional[dict] = None) -> torch.Tensor:
    """uint8 NCHW images -> stem convolution output (bf16, channels-last), see :class:`_StemS2DFn`.  ``bn``: statistics
    buffers of the BatchNorm that follows (ops/bn.py ``stats_buffers``), filled by the convolution's epilogue on the
    tensor-core path (check :func:`stem_stats_fused`)."""
    N, _, H, W = img.shape
    xs = torch.empty((N, 16, H // 2 + 3, W // 2 + 3), device=img.device, dtype=torch.bfloat16, memory_format=torch.channels_last)
    count(1)
    native().image_normalize_s2d(img.data_ptr(), xs.data_ptr(), N, H, W, float(mean[0]), float(mean[1]), float(mean[2]),
                                 1.0 / float(std[0]), 1.0 / float(std[1]), 1.0 / float(std[2]), stream_ptr())
    return _StemS2DFn.apply(xs, conv.weight, getattr(conv, "w_bf16", None), bn if stem_stats_fused(conv) else None)


def stem_stats_fused(conv: nn.Conv2d) -> bool:
    """The stem runs on the tcgen05 kernel (which can produce the BatchNorm statistics in its epilogue)."""
    return _stem_tc() and conv.weight.shape[0] % 64 == 0
