"""Convolution with an fp32 master filter and a bf16 compute copy (ResNet-50 local step).

The convolution itself is library code (cuDNN's sm_100 implicit-GEMM kernels through
``aten::convolution`` -- the same role cuBLAS plays for plain GEMMs); what this layer removes is
everything *around* it that the autocast path launches per layer and per step
(profiles/launches_resnet50_fusedbn_v2_r1.txt):

* the fp32 -> bf16 filter cast: the forward reads the bf16 shadow copy that the fused optimizer (K7)
  and the aggregation kernel (K2) keep up to date,
* the NCHW -> NHWC filter conversion: filters live channels-last inside the flat buffers
  (models/flat.py),
* the bf16 -> fp32 gradient cast and the gradient accumulation: the bf16 filter gradient is queued on
  the flat model's gradient sink and added into the flat fp32 gradient buffer by ONE multi-tensor
  kernel per step (ops/optim.py::multi_accumulate).

Without a shadow (CPU tests, the stock-optimizer baseline arm) it is a plain ``nn.Conv2d``.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
from torch import nn


class _ShadowConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, w_bf16, stride, padding, sink, offset):
        y = torch.ops.aten.convolution(x, w_bf16, None, stride, padding, (1, 1), False, (0, 0), 1)
        ctx.save_for_backward(x, w_bf16)
        ctx.conf = (stride, padding)
        ctx.sink, ctx.offset = sink, offset
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w_bf16 = ctx.saved_tensors
        stride, padding = ctx.conf
        dx, dw, _ = torch.ops.aten.convolution_backward(dy, x, w_bf16, None, stride, padding, (1, 1), False, (0, 0), 1,
                                                        (ctx.needs_input_grad[0], True, False))
        if not dw.is_contiguous(memory_format=torch.channels_last):
            dw = dw.contiguous(memory_format=torch.channels_last)
        ctx.sink.append((dw, ctx.offset))        # flushed by FlatModel.flush_grad_sink() after backward
        return dx, None, None, None, None, None, None


class ShadowConv2d(nn.Conv2d):
    """``nn.Conv2d`` (no bias, groups=1, dilation=1) whose forward uses the bf16 shadow filter when attached."""

    def __init__(self, *args, **kw):
        super().__init__(*args, **kw)
        assert self.bias is None and self.groups == 1 and tuple(self.dilation) == (1, 1) and self.padding_mode == "zeros"
        self.w_bf16: Optional[torch.Tensor] = None        # channels-last view into the bf16 shadow buffer
        self._sink: Optional[List[Tuple[torch.Tensor, int]]] = None
        self._offset = 0

    def attach(self, w_bf16: torch.Tensor, sink: list, offset: int) -> None:
        self.w_bf16, self._sink, self._offset = w_bf16, sink, int(offset)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.w_bf16 is None or not x.is_cuda or not torch.is_grad_enabled() or not self.weight.requires_grad:
            if self.w_bf16 is not None and x.is_cuda:
                return torch.ops.aten.convolution(x.to(torch.bfloat16), self.w_bf16, None, self.stride, self.padding, (1, 1),
                                                  False, (0, 0), 1)
            return super().forward(x)
        if x.dtype != torch.bfloat16:
            x = x.to(torch.bfloat16)
        return _ShadowConvFn.apply(x, self.weight, self.w_bf16, tuple(self.stride), tuple(self.padding), self._sink,
                                   self._offset)
