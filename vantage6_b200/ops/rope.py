"""K6: rotary position embedding for Llama (rotate-half), in place on q and k in one launch."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import native, stream_ptr


def rope_tables(seq_len: int, head_dim: int, theta: float = 500000.0, device="cpu") -> Tuple[torch.Tensor, torch.Tensor]:
    """fp32 cos/sin tables of shape [seq_len, head_dim/2] (Llama-3 default theta = 5e5)."""
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float64) / head_dim))
    ang = torch.arange(seq_len, dtype=torch.float64)[:, None] * inv[None, :]
    return ang.cos().float().to(device).contiguous(), ang.sin().float().to(device).contiguous()


class _RopeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, cos, sin, pos_ids):
        assert q.dtype == torch.bfloat16
        # rotate copies: q/k are usually views of a projection output, and autograd does not allow a
        # multi-output Function to modify views in place
        q, k = q.contiguous().clone(), k.contiguous().clone()
        B, S, Hq, D = q.shape
        Hkv = k.shape[2]
        native().rope(q.data_ptr(), k.data_ptr(), cos.data_ptr(), sin.data_ptr(),
                      0 if pos_ids is None else pos_ids.data_ptr(), B, S, Hq, Hkv, D, False, stream_ptr())
        ctx.save_for_backward(cos, sin, pos_ids if pos_ids is not None else torch.empty(0))
        ctx.has_pos = pos_ids is not None
        return q, k

    @staticmethod
    def backward(ctx, dq, dk):
        cos, sin, pos = ctx.saved_tensors
        dq = dq.contiguous().clone()
        dk = dk.contiguous().clone()
        B, S, Hq, D = dq.shape
        native().rope(dq.data_ptr(), dk.data_ptr(), cos.data_ptr(), sin.data_ptr(),
                      pos.data_ptr() if ctx.has_pos else 0, B, S, Hq, dk.shape[2], D, True, stream_ptr())
        return dq, dk, None, None, None


def apply_rope(q: torch.Tensor, k: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor,
               pos_ids: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """q:[B,S,Hq,D], k:[B,S,Hkv,D] bf16 -> rotated (in place on CUDA)."""
    if q.is_cuda:
        return _RopeFn.apply(q, k, cos, sin, pos_ids)
    return reference_rope(q, k, cos, sin, pos_ids)


def reference_rope(q, k, cos, sin, pos_ids=None):
    def rot(x):
        B, S, H, D = x.shape
        c = cos[:S] if pos_ids is None else cos[pos_ids]
        s = sin[:S] if pos_ids is None else sin[pos_ids]
        c = c.reshape(-1, S, 1, D // 2) if pos_ids is not None else c[None, :, None, :]
        s = s.reshape(-1, S, 1, D // 2) if pos_ids is not None else s[None, :, None, :]
        x1, x2 = x.float()[..., : D // 2], x.float()[..., D // 2:]
        return torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], dim=-1).to(x.dtype)

    return rot(q), rot(k)
