"""Convolution on the hand-written tcgen05 implicit-GEMM kernels (csrc/igemm.cu).

Three entry points, all on NHWC bf16 activations and ``[Cout, R, S, Cin]`` bf16 filters (a PyTorch
``channels_last`` tensor *is* that memory layout, so no conversion happens anywhere):

* :func:`conv_fprop`  -- ``y = conv(x, w)``; with ``bn=...`` the BatchNorm batch statistics of ``y`` (mean, rstd, the
  fused scale/bias, the running-stat update) come out of the same launch, so the separate statistics pass over ``y``
  disappears (ops/bn.py then only runs the apply pass),
* :func:`conv_dgrad`  -- ``dx = conv_transpose(dy, w)`` (stride 1; the filter is read in place through an MN-major
  UMMA descriptor, no transposed / flipped copy),
* :func:`conv_wgrad`  -- ``dw += dy^T . im2col(x)`` accumulated in fp32 **directly into the flat gradient buffer**
  (split-K over pixels, ``red.global.add.v4.f32``), so the bf16 gradient + gradient-sink round trip disappears.

``TcConv2d`` packages them as an autograd function; ``supported()`` says which layer shapes the kernels take
(Cin, Cout multiples of 64, square filter, stride-1 data gradient) -- everything in ResNet-50 except the 16-channel
space-to-depth stem and the data gradient of the six stride-2 layers.

Reference parity: the reference has no compute; the contract is BASELINE.json config 2.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch

from . import count, native, stream_ptr

_scratch: dict = {}


def igemm_scratch(device) -> torch.Tensor:
    """Counters + per-CTA statistics partials of the FPROP statistics epilogue (zeroed once; the counters reset
    themselves; launches are stream-ordered)."""
    key = str(device)
    buf = _scratch.get(key)
    if buf is None:
        buf = _scratch[key] = torch.zeros(int(native().IGEMM_SCRATCH_FLOATS), device=device, dtype=torch.float32)
    return buf


def _nhwc(x: torch.Tensor) -> Tuple[int, int, int, int]:
    n, c, h, w = x.shape
    return n, h, w, c


def _is_cl(x: torch.Tensor) -> bool:
    return x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last)


def supported(cin: int, cout: int, r: int, s: int, stride: int, pad: int) -> bool:
    """Layer shapes the tcgen05 kernels take (forward + filter gradient; the data gradient additionally needs stride 1)."""
    return cin % 64 == 0 and cout % 64 == 0 and r == s and stride in (1, 2) and 0 <= pad < r


def dgrad_supported(cin: int, cout: int, r: int, s: int, stride: int, pad: int) -> bool:
    return supported(cin, cout, r, s, stride, pad) and stride == 1


def out_size(h: int, r: int, stride: int, pad: int) -> int:
    return (h + 2 * pad - r) // stride + 1


def conv_fprop(x: torch.Tensor, w: torch.Tensor, stride: int = 1, pad: int = 0, bn: Optional[dict] = None,
               bias: Optional[torch.Tensor] = None, act: int = 0, force_im2col: bool = False) -> torch.Tensor:
    """``x``: [N, Cin, H, W] channels_last bf16; ``w``: [Cout, Cin, R, S] channels_last bf16 -> y channels_last bf16.

    ``bn``: dict(gamma, beta, running_mean, running_var, num_batches_tracked, mean, rstd, scale_bias, eps, momentum);
    mean / rstd / scale_bias ([2*Cout]) are outputs."""
    assert x.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and _is_cl(x) and _is_cl(w)
    n, h, wd, cin = _nhwc(x)
    cout, cin2, r, s = w.shape
    assert cin == cin2
    p, q = out_size(h, r, stride, pad), out_size(wd, s, stride, pad)
    y = torch.empty((n, cout, p, q), device=x.device, dtype=torch.bfloat16, memory_format=torch.channels_last)
    g = bn or {}

    def ptr(t):
        return 0 if t is None else t.data_ptr()

    count(1)
    native().conv_fprop(x.data_ptr(), w.data_ptr(), y.data_ptr(), ptr(bias), act, n, h, wd, cin, cout, r, s, stride, pad,
                        ptr(g.get("gamma")), ptr(g.get("beta")), ptr(g.get("running_mean")), ptr(g.get("running_var")),
                        ptr(g.get("num_batches_tracked")), ptr(g.get("mean")), ptr(g.get("rstd")), ptr(g.get("scale_bias")),
                        igemm_scratch(x.device).data_ptr() if bn else 0, float(g.get("eps", 1e-5)), float(g.get("momentum", 0.1)),
                        force_im2col, stream_ptr())
    return y


def conv_dgrad(dy: torch.Tensor, w: torch.Tensor, in_hw: Tuple[int, int], pad: int = 0, force_im2col: bool = False) -> torch.Tensor:
    """Data gradient of a stride-1 convolution: ``dy`` [N, Cout, P, Q] -> dx [N, Cin, H, W] (channels_last bf16)."""
    assert dy.is_cuda and dy.dtype == torch.bfloat16 and _is_cl(dy) and _is_cl(w)
    n, p, q, cout = _nhwc(dy)
    cout2, cin, r, s = w.shape
    h, wd = in_hw
    assert cout == cout2 and p == h + 2 * pad - r + 1 and q == wd + 2 * pad - s + 1
    dx = torch.empty((n, cin, h, wd), device=dy.device, dtype=torch.bfloat16, memory_format=torch.channels_last)
    count(1)
    native().conv_dgrad(dy.data_ptr(), w.data_ptr(), dx.data_ptr(), n, h, wd, cin, cout, r, s, pad, force_im2col, stream_ptr())
    return dx


def conv_wgrad(dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor, rs: Tuple[int, int], stride: int = 1, pad: int = 0,
               scale: float = 1.0, splits: int = 0, force_im2col: bool = False) -> torch.Tensor:
    """``dw`` (fp32, memory [Cout, R, S, Cin], e.g. the ``.grad`` view of a flat model) ``+= scale * dy^T . im2col(x)``."""
    assert dy.is_cuda and dy.dtype == torch.bfloat16 and x.dtype == torch.bfloat16 and _is_cl(dy) and _is_cl(x)
    assert dw.dtype == torch.float32
    n, h, wd, cin = _nhwc(x)
    cout = dy.shape[1]
    r, s = rs
    assert dw.numel() == cout * r * s * cin
    count(1)
    native().conv_wgrad(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), n, h, wd, cin, cout, r, s, stride, pad, float(scale), int(splits),
                        force_im2col, stream_ptr())
    return dw


FORCE_IM2COL = os.environ.get("V6B200_CONV_FORCE_IM2COL") == "1"
