"""K5: LayerNorm / RMSNorm forward+backward with fused residual add (csrc/norm.cu).

``layer_norm(x, gamma, beta, eps, residual=None)`` returns ``y`` (and the updated residual
stream when ``residual`` is given).  gamma/beta are fp32 (views into the flat master buffer);
activations are bf16 or fp32.  Autograd is wired through ``torch.autograd.Function`` so the
hand-written backward kernel is what runs in training.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import native, stream_ptr

_SCRATCH_PARTS = 148 * 2
_scratch: dict = {}


def _get_scratch(device, cols: int) -> torch.Tensor:
    key = (device, cols)
    t = _scratch.get(key)
    if t is None:
        t = torch.empty(2 * _SCRATCH_PARTS * cols, device=device, dtype=torch.float32)
        _scratch[key] = t
    return t


def _ptr(t):
    return 0 if t is None else t.data_ptr()


class _LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, residual, eps, rms):
        shape = x.shape
        cols = shape[-1]
        x2 = x.reshape(-1, cols).contiguous()
        rows = x2.shape[0]
        res2 = residual.reshape(-1, cols).contiguous() if residual is not None else None
        y = torch.empty_like(x2)
        res_out = torch.empty_like(x2) if residual is not None else None
        rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
        mean = None if rms else torch.empty(rows, device=x.device, dtype=torch.float32)
        bf16 = x.dtype == torch.bfloat16
        C = native()
        if rms:
            C.rmsnorm_fwd(x2.data_ptr(), _ptr(res2), gamma.data_ptr(), y.data_ptr(), _ptr(res_out), rstd.data_ptr(),
                          rows, cols, eps, bf16, stream_ptr())
        else:
            C.layernorm_fwd(x2.data_ptr(), _ptr(res2), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), _ptr(res_out),
                            mean.data_ptr(), rstd.data_ptr(), rows, cols, eps, bf16, stream_ptr())
        x_in = res_out if residual is not None else x2
        ctx.save_for_backward(x_in, gamma, mean if mean is not None else rstd, rstd)
        ctx.rms = rms
        ctx.has_res = residual is not None
        ctx.shape = shape
        ctx.has_beta = beta is not None
        ctx.params = (gamma, beta)                # the Parameters themselves: .grad may be a flat-buffer view
        if residual is not None:
            return y.view(shape), res_out.view(shape)
        return y.view(shape), None

    @staticmethod
    def backward(ctx, dy, dres):
        x_in, gamma, mean, rstd = ctx.saved_tensors
        cols = x_in.shape[-1]
        rows = x_in.shape[0]
        dy2 = dy.reshape(-1, cols).contiguous()
        dres2 = dres.reshape(-1, cols).contiguous() if (ctx.has_res and dres is not None) else None
        dx = torch.empty_like(x_in)
        # flat-buffer models pre-allocate .grad as views of one fp32 buffer: accumulate dgamma / dbeta in place
        pg, pb = ctx.params
        direct = all(p is None or (p.grad is not None and p.grad.dtype == torch.float32 and p.grad.is_contiguous())
                     for p in ((pg,) if ctx.rms else (pg, pb))) and pg is not None and (ctx.rms or pb is not None)
        frozen = pg is not None and not pg.requires_grad and (ctx.rms or pb is None or not pb.requires_grad)
        dgamma = None if frozen else (pg.grad if direct else torch.empty_like(gamma))
        dbeta = None if (ctx.rms or frozen) else (pb.grad if direct else torch.empty_like(gamma))
        scratch = _get_scratch(x_in.device, cols)
        bf16 = x_in.dtype == torch.bfloat16
        C = native()
        if ctx.rms:
            C.rmsnorm_bwd(dy2.data_ptr(), x_in.data_ptr(), _ptr(dres2), gamma.data_ptr(), rstd.data_ptr(), dx.data_ptr(),
                          _ptr(dgamma), scratch.data_ptr(), _SCRATCH_PARTS, rows, cols, direct, bf16, stream_ptr())
        else:
            C.layernorm_bwd(dy2.data_ptr(), x_in.data_ptr(), _ptr(dres2), gamma.data_ptr(), mean.data_ptr(),
                            rstd.data_ptr(), dx.data_ptr(), _ptr(dgamma), _ptr(dbeta), scratch.data_ptr(),
                            _SCRATCH_PARTS, rows, cols, direct, bf16, stream_ptr())
        dxv = dx.view(ctx.shape)
        if direct or frozen:
            return dxv, None, None, (dxv if ctx.has_res else None), None, None
        return dxv, dgamma, (dbeta if ctx.has_beta else None), (dxv if ctx.has_res else None), None, None


def layer_norm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5,
               residual: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """y = LN(x + residual) * gamma + beta; returns (y, x + residual or None)."""
    if x.is_cuda:
        return _LayerNormFn.apply(x, gamma, beta, residual, eps, False)
    return reference_layer_norm(x, gamma, beta, eps, residual)


def rms_norm(x: torch.Tensor, gamma: torch.Tensor, eps: float = 1e-5,
             residual: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    if x.is_cuda:
        return _LayerNormFn.apply(x, gamma, None, residual, eps, True)
    return reference_rms_norm(x, gamma, eps, residual)


# ------------------------------------------------------------------ references (fp32 math)
def reference_layer_norm(x, gamma, beta, eps=1e-5, residual=None):
    h = x.float() + (residual.float() if residual is not None else 0.0)
    y = torch.nn.functional.layer_norm(h, (h.shape[-1],), gamma.float(), beta.float(), eps)
    return y.to(x.dtype), (h.to(x.dtype) if residual is not None else None)


def reference_rms_norm(x, gamma, eps=1e-5, residual=None):
    h = x.float() + (residual.float() if residual is not None else 0.0)
    y = h * torch.rsqrt(h.pow(2).mean(-1, keepdim=True) + eps) * gamma.float()
    return y.to(x.dtype), (h.to(x.dtype) if residual is not None else None)
