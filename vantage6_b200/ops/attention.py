"""K4: flash attention.  Forward = hand-written tcgen05/TMEM/TMA kernels: ``csrc/attention.cu`` (one 128-query
tile per SM, intra-CTA pipeline) and ``csrc/attention2.cu`` (two co-resident CTAs per SM, P aliased over S in
TMEM); ``V6B200_ATTN_FWD=1cta|2cta`` or the ``variant`` argument selects one.  Backward = the flash-attn library
kernel fed with our output and log-sum-exp by default (it plays the role cuBLAS plays for dX/dW), or the
hand-written transpose-free tcgen05 backward (``csrc/attention_bwd.cu``, ``V6B200_ATTN_BWD=native``).

q:[B,S,Hq,D]  k,v:[B,S,Hkv,D]  bf16, D in {64,128}, GQA allowed -> o:[B,S,Hq,D].
"""
from __future__ import annotations

import math
import os

import torch

from . import count, native, stream_ptr

_enabled = os.environ.get("V6B200_ATTENTION", "1") != "0"


def available() -> bool:
    """True when the tcgen05 forward can be used (GPU + extension + not disabled by env)."""
    return _enabled and torch.cuda.is_available() and native(required=False) is not None \
        and hasattr(native(), "flash_attn_fwd")


def _supported(q: torch.Tensor, k: torch.Tensor) -> bool:
    D, S = q.shape[-1], q.shape[1]
    return q.dtype == torch.bfloat16 and D in (64, 128) and S % 8 == 0 and q.shape[2] % k.shape[2] == 0


def _fwd_variant(variant: str | None) -> str:
    v = variant or os.environ.get("V6B200_ATTN_FWD", _DEFAULT_FWD)
    return v if v in ("1cta", "2cta") and (v == "1cta" or hasattr(native(), "flash_attn_fwd2")) else "1cta"


_DEFAULT_FWD = "2cta"      # measured: 1.1-1.7x the 1-CTA design and 1.17-1.74x flash-attn 2 (profiles/kernel_bench_attn_r1c.json)


def flash_attn_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool, scale: float | None = None,
                   variant: str | None = None):
    """Raw forward: returns (o, lse).  V is transposed to [B,Hkv,D,S] so the PV operand is K-major.  q / k may be
    token-strided views (e.g. slices of a packed QKV projection): the TMA tensor maps take the row stride."""
    B, S, Hq, D = q.shape
    Hkv = k.shape[2]
    scale = scale if scale is not None else 1.0 / math.sqrt(D)
    q, ldq = _token_strided(q)
    k, ldk = _token_strided(k)
    o = torch.empty((B, S, Hq, D), device=q.device, dtype=q.dtype)
    lse = torch.empty(B, Hq, S, device=q.device, dtype=torch.float32)
    count(1)
    if os.environ.get("V6B200_ATTN_V", "mn") == "mn" and _fwd_variant(variant) == "2cta" and hasattr(native(), "flash_attn_fwd2_vmn"):
        # default: V consumed in place as an MN-major UMMA operand -- no Vt copy (validated on hardware in round 2: S = 128:
        # 28.7 vs 45.3 us, causal 2048 x 128: 0.350 vs 0.386 ms, BERT-base round 27.6 vs 28.3 ms; V6B200_ATTN_V=t: transposed copy)
        v, ldv = _token_strided(v)
        native().flash_attn_fwd2_vmn(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), lse.data_ptr(), B, S, Hq, Hkv, D,
                                     ldq, ldk, ldv, float(scale), bool(causal), stream_ptr())
        return o, lse
    vt = v.permute(0, 2, 3, 1).contiguous()
    fn = native().flash_attn_fwd2 if _fwd_variant(variant) == "2cta" else native().flash_attn_fwd
    fn(q.data_ptr(), k.data_ptr(), vt.data_ptr(), o.data_ptr(), lse.data_ptr(), B, S, Hq, Hkv, D, ldq, ldk, float(scale),
       bool(causal), stream_ptr())
    return o, lse


def _token_strided(t: torch.Tensor):
    """(tensor, row stride in elements) if ``t`` [B,S,H,D] is dense in (H,D) with one uniform token stride; else a
    contiguous copy."""
    B, S, H, D = t.shape
    st = t.stride()
    if st[3] == 1 and st[2] == D and st[1] % 8 == 0 and st[1] >= H * D and (B == 1 or st[0] == S * st[1]) \
            and t.data_ptr() % 16 == 0:
        return t, st[1]
    t = t.contiguous()
    return t, H * D


def _bwd_native() -> bool:
    """Hand-written tcgen05 backward (csrc/attention_bwd.cu) vs the flash-attn library backward."""
    return os.environ.get("V6B200_ATTN_BWD", "lib") == "native" and hasattr(native(), "flash_attn_bwd")


def flash_attn_bwd(do: torch.Tensor, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, o: torch.Tensor, lse: torch.Tensor,
                   causal: bool, scale: float | None = None):
    """dQ, dK, dV on the tcgen05 kernels.  Pre-pass (plain torch ops): delta = rowsum(dO o O), lse in
    log2 units, and the [B,H,D,S] transposes that make every MMA B operand K-major."""
    B, S, Hq, D = q.shape
    Hkv = k.shape[2]
    scale = scale if scale is not None else 1.0 / math.sqrt(D)
    do = do.contiguous()
    delta = (do.float() * o.float()).sum(-1).permute(0, 2, 1).contiguous()            # [B,Hq,S]
    lse2 = (lse * 1.4426950408889634).contiguous()
    kt = k.permute(0, 2, 3, 1).contiguous()
    qt = q.permute(0, 2, 3, 1).contiguous()
    dot = do.permute(0, 2, 3, 1).contiguous()
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    count(2)
    native().flash_attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), do.data_ptr(), kt.data_ptr(), qt.data_ptr(),
                            dot.data_ptr(), lse2.data_ptr(), delta.data_ptr(), dq.data_ptr(), dk.data_ptr(), dv.data_ptr(),
                            B, S, Hq, Hkv, D, float(scale), bool(causal), stream_ptr())
    return dq, dk, dv


class _FlashAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, causal):
        scale = 1.0 / math.sqrt(q.shape[-1])
        o, lse = flash_attn_fwd(q, k, v, causal, scale)        # strided q / k are consumed in place
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.causal, ctx.scale = causal, scale
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o, lse = ctx.saved_tensors
        if _bwd_native():
            dq, dk, dv = flash_attn_bwd(do, q, k, v, o, lse, ctx.causal, ctx.scale)
            return dq, dk, dv, None
        from flash_attn.flash_attn_interface import _flash_attn_backward

        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        _flash_attn_backward(do.contiguous(), q, k, v, o, lse, dq, dk, dv, 0.0, ctx.scale, ctx.causal, -1, -1, 0.0, None,
                             False, None)
        return dq, dk, dv, None


class _PackedQKVAttnFn(torch.autograd.Function):
    """Self-attention on a packed projection output qkv:[B,S,3,H,D] (MHA): q / k are read in place through strided
    tensor maps, and the backward writes dq / dk / dv straight into one dqkv buffer -- no slice copies either way."""

    @staticmethod
    def forward(ctx, qkv, causal):
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
        scale = 1.0 / math.sqrt(q.shape[-1])
        o, lse = flash_attn_fwd(q, k, v, causal, scale)
        ctx.save_for_backward(qkv, o, lse)
        ctx.causal, ctx.scale = causal, scale
        return o

    @staticmethod
    def backward(ctx, do):
        qkv, o, lse = ctx.saved_tensors
        q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
        dqkv = torch.empty_like(qkv)
        if _bwd_native():
            dq, dk, dv = flash_attn_bwd(do, q.contiguous(), k.contiguous(), v.contiguous(), o, lse, ctx.causal, ctx.scale)
            dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2] = dq, dk, dv
            return dqkv, None
        from flash_attn.flash_attn_interface import _flash_attn_backward

        _flash_attn_backward(do.contiguous(), q, k, v, o, lse, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2], 0.0, ctx.scale,
                             ctx.causal, -1, -1, 0.0, None, False, None)
        return dqkv, None


def packed_qkv_attention(qkv: torch.Tensor, causal: bool = False) -> torch.Tensor:
    """qkv:[B,S,3,H,D] bf16 -> o:[B,S,H,D]."""
    return _PackedQKVAttnFn.apply(qkv, causal)


def flash_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool = False) -> torch.Tensor:
    if not _supported(q, k):
        raise ValueError(f"unsupported attention shape/dtype for the tcgen05 kernel: q={tuple(q.shape)} {q.dtype}")
    return _FlashAttnFn.apply(q, k, v, causal)


def reference_attention(q, k, v, causal=False):
    """fp32 PyTorch reference (tests)."""
    B, S, Hq, D = q.shape
    Hkv = k.shape[2]
    qf, kf, vf = (t.float().transpose(1, 2) for t in (q, k, v))
    if Hkv != Hq:
        kf = kf.repeat_interleave(Hq // Hkv, dim=1)
        vf = vf.repeat_interleave(Hq // Hkv, dim=1)
    s = qf @ kf.transpose(-1, -2) / math.sqrt(D)
    if causal:
        s = s.masked_fill(torch.ones(S, S, device=q.device, dtype=torch.bool).triu(1), float("-inf"))
    p = torch.softmax(s, dim=-1)
    return (p @ vf).transpose(1, 2), torch.logsumexp(s, dim=-1)
