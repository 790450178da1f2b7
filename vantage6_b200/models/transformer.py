"""Transformer building blocks wired to the hand-written sm_100a ops.

Parameter policy (mixed precision without autocast):
* trainable parameters are fp32 and live in the flat master buffer (models/flat.py);
* GEMM weights are *consumed* in bf16: either from the bf16 shadow buffer that the fused
  optimizer (K7) and the aggregation kernel (K2) keep up to date -- no cast kernel in the step --
  or from an on-the-fly cast when no shadow is attached (CPU / tests);
* activations are bf16; norms read fp32 gamma/beta directly from the master buffer.

``ShadowLinear`` runs its forward on the tcgen05 GEMM (ops/gemm.py); backward GEMMs (dX, dW) are
plain library GEMMs (cuBLAS); the activation derivative + bias gradient is one hand-written kernel
(csrc/act.cu) and the bf16 dW of every layer is added into the flat fp32 gradient buffer by the
multi-tensor gradient sink (one launch per step).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
from torch import nn

from ..ops import gemm as G
from ..ops import norm as N


def _tc_linear_bwd(dy2: torch.Tensor, w: torch.Tensor) -> bool:
    """Backward GEMMs of the linear layers on the hand-written tcgen05 kernels (ops/conv.py::linear_dgrad /
    linear_wgrad, the 1x1 case of the implicit-GEMM family): default on; ``V6B200_LINEAR_BWD=cublas`` selects the library."""
    import os

    from ..ops import conv as C

    return (dy2.is_cuda and dy2.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and w.is_contiguous()
            and os.environ.get("V6B200_LINEAR_BWD", "tc") == "tc" and C.linear_bwd_supported(w.shape[0], w.shape[1]))


def _mm_f32(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """bf16 x bf16 -> fp32 (cuBLAS, fp32 output when the build supports it)."""
    try:
        return torch.mm(a, b, out_dtype=torch.float32)
    except TypeError:
        return torch.mm(a, b).float()
    except RuntimeError:
        return torch.mm(a, b).float()


# K1: in the FIRST local step of a round the layers tagged ``first_consumer`` do not find their new bf16
# weights in the local shadow buffer -- the aggregation kernel left them on their owner (FedAvgEngine ``shadow_skip``) --
# but receive them inside their forward GEMM: the owner's kernel multicasts the tiles through the NVSwitch, every rank's
# main loop consumes them behind per-tile flags (ops/gemm.py::bcast_push_gemm_bf16).  The trainer switches this on for
# that one step (``K1_STEP["on"]``); every other step is the plain GEMM on the (by then complete) local copy.
K1_STEP = {"on": False}


class _ShadowLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, w_bf16, act, sink, offset, k1=None):
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        wb = w_bf16 if w_bf16 is not None else weight.detach().to(torch.bfloat16)
        if x2.is_cuda and k1 is not None:
            pre = G.bcast_push_gemm_bf16(x2, wb, k1["w_mc_ptr"], k1["flags"], k1["flag_peer_ptrs"], k1["world"], k1["is_owner"], 0,
                                         bias=bias, own_blocks=k1["own_blocks"], epoch_ptr=k1["epoch_ptr"], status_ptr=k1["status_ptr"])
        elif x2.is_cuda:
            pre = G.gemm_bf16(x2, wb, bias, G.ACT_NONE)
        else:
            pre = G.reference_linear(x2, wb, bias, G.ACT_NONE)
        y = pre
        if act == G.ACT_GELU:
            y = torch.nn.functional.gelu(pre)
        elif act == G.ACT_RELU:
            y = torch.relu(pre)
        ctx.save_for_backward(x2, wb, pre if act != G.ACT_NONE else None)
        ctx.act, ctx.xshape = act, x.shape
        ctx.need_w = weight.requires_grad
        ctx.weight = weight
        ctx.bias = bias                           # the Parameter itself: its .grad may be a flat-buffer view
        ctx.sink, ctx.offset = sink, offset
        return y.view(*x.shape[:-1], wb.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, wb, pre = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        # one kernel: activation derivative, bf16 dpre, bias gradient (accumulated in place when possible)
        dy2, db = G.bias_act_backward(dy2, pre, ctx.act, ctx.bias)
        tc = _tc_linear_bwd(dy2, wb)
        if tc:
            from ..ops import conv as C

            dx = C.linear_dgrad(dy2, wb).view(ctx.xshape) if ctx.needs_input_grad[0] else None
            g = ctx.weight.grad if ctx.need_w else None
            if g is not None and g.dtype == torch.float32 and g.is_contiguous() and g.shape == wb.shape:
                C.linear_wgrad(dy2, x2, g)              # fp32, straight into the flat gradient buffer: no bf16 dW, no sink entry
                return dx, None, db, None, None, None, None, None
            if ctx.need_w:
                dwf = torch.zeros(wb.shape, device=dy2.device, dtype=torch.float32)
                C.linear_wgrad(dy2, x2, dwf)
                return dx, dwf, db, None, None, None, None, None
            return dx, None, db, None, None, None, None, None
        dx = torch.mm(dy2, wb).view(ctx.xshape) if ctx.needs_input_grad[0] else None
        dw = None
        if ctx.need_w:
            if ctx.sink is not None and dy2.is_cuda:
                ctx.sink.append((torch.mm(dy2.t(), x2), ctx.offset))      # bf16 dW -> flat fp32 grads, one kernel per step
            else:
                dw = _mm_f32(dy2.t(), x2)
        return dx, dw, db, None, None, None, None, None


class ShadowLinear(nn.Module):
    """nn.Linear with fp32 master weight, bf16 compute copy and optional fused activation."""

    def __init__(self, in_features: int, out_features: int, bias: bool = True, act: int = G.ACT_NONE,
                 init_std: float = 0.02, first_consumer: bool = False):
        super().__init__()
        self.in_features, self.out_features, self.act = in_features, out_features, act
        self.first_consumer = bool(first_consumer)      # first GEMM of its block to read the new global weights (K1 candidate)
        self.k1: Optional[dict] = None                  # set by the trainer (bcast="fused"): FedAvgEngine.k1_layer(...)
        self.weight = nn.Parameter(torch.empty(out_features, in_features).normal_(0.0, init_std))
        self.bias = nn.Parameter(torch.zeros(out_features)) if bias else None
        self.w_bf16: Optional[torch.Tensor] = None      # view into the shadow buffer (set by attach_shadow)
        self._sink = None                               # the flat model's gradient sink (set by attach_shadow)
        self._offset = 0

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dtype != torch.bfloat16:
            x = x.to(torch.bfloat16)
        k1 = self.k1 if (K1_STEP["on"] and self.k1 is not None) else None
        return _ShadowLinearFn.apply(x, self.weight, self.bias, self.w_bf16, self.act, self._sink, self._offset, k1)


class FrozenLinear(nn.Module):
    """Frozen bf16 base weight (Llama LoRA): forward and dX on the tcgen05 GEMM (dX through a transposed copy made once,
    :func:`frozen_transposed`); the weight is a buffer-less plain tensor so it is neither federated nor optimised."""

    def __init__(self, in_features: int, out_features: int, device=None, init_std: float = 0.02):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        w = torch.empty(out_features, in_features, dtype=torch.bfloat16, device=device)
        w.normal_(0.0, init_std)
        self.weight_bf16 = w                      # deliberately not a Parameter / buffer

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        self.weight_bf16 = fn(self.weight_bf16)
        return self

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _FrozenLinearFn.apply(x, self.weight_bf16, self)


def frozen_transposed(holder) -> Optional[torch.Tensor]:
    """[K_in, N_out] copy of the FROZEN bf16 weight of ``holder`` (a :class:`FrozenLinear`), made once and owned by the module:
    dX = dY . W then runs on the K-major tcgen05 GEMM of ops/gemm.py (0.88-0.99x cuBLAS on the Llama shapes) instead of the
    MN-major implicit-GEMM form or cuBLAS.  Costs one extra copy of the frozen weights (Llama-3 8B: 15 GB of 180);
    ``V6B200_FROZEN_DX=igemm|cublas`` selects the other paths.  Re-made when the weight was moved or overwritten in place."""
    import os

    w = getattr(holder, "weight_bf16", None)
    if (w is None or os.environ.get("V6B200_FROZEN_DX", "gemm") != "gemm" or not w.is_cuda or w.dtype != torch.bfloat16 or w.dim() != 2
            or w.shape[0] % 8 or w.shape[1] % 8):
        return None
    tag = (w.data_ptr(), w._version, tuple(w.shape))
    if getattr(holder, "_wt_tag", None) != tag:
        holder._wt = w.t().contiguous()
        holder._wt_tag = tag
    return holder._wt


class _FrozenLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, holder=None):
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        y = G.gemm_bf16(x2, w) if x2.is_cuda else G.reference_linear(x2, w)
        ctx.save_for_backward(w)
        ctx.xshape = x.shape
        ctx.holder = holder
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        (w,) = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        wt = frozen_transposed(ctx.holder) if (dy2.dtype == torch.bfloat16 and ctx.holder is not None) else None
        if wt is not None:
            return G.gemm_bf16(dy2, wt).view(ctx.xshape), None, None
        if _tc_linear_bwd(dy2, w):
            from ..ops import conv as C

            return C.linear_dgrad(dy2, w).view(ctx.xshape), None, None
        return torch.mm(dy2, w).view(ctx.xshape), None, None


class _LoRALinearFn(torch.autograd.Function):
    """y = x W^T + s (x A^T) B^T as ONE autograd node: 3 launches forward (tcgen05 base GEMM, x A^T, fused
    ``addmm``), 6 backward, no per-call casts when the bf16 shadows of A / B are attached, and the adapter gradients
    go through the multi-tensor gradient sink.  (The composed PyTorch expression cost ~20 small launches per adapter
    and step -- 128 adapters in Llama-3-8B.)"""

    @staticmethod
    def forward(ctx, x, lora_a, lora_b, w_base, a16, b16, scaling, sink, off_a, off_b, holder=None):
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        cdt = x2.dtype
        a_w = a16 if a16 is not None else lora_a.detach().to(cdt)            # [r, K]
        b_w = b16 if b16 is not None else lora_b.detach().to(cdt)            # [N, r]
        y = G.gemm_bf16(x2, w_base) if x2.is_cuda else G.reference_linear(x2, w_base.to(cdt))
        xa_s = torch.mm(x2, a_w.t()) * scaling                               # [M, r], scaling folded in once
        y = y.addmm_(xa_s, b_w.t()) if y.dtype == xa_s.dtype else y + (xa_s @ b_w.t()).to(y.dtype)
        ctx.save_for_backward(x2, xa_s, a_w, b_w, w_base)
        ctx.scaling, ctx.xshape = scaling, x.shape
        ctx.sink, ctx.off_a, ctx.off_b = sink, off_a, off_b
        ctx.need_ab = lora_a.requires_grad
        ctx.holder = holder
        return y.view(*x.shape[:-1], w_base.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, xa_s, a_w, b_w, w_base = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        cdt = x2.dtype
        if dy2.dtype != cdt:
            dy2 = dy2.to(cdt)
        da_s = torch.mm(dy2, b_w) * ctx.scaling                              # [M, r]
        dx = None
        if ctx.needs_input_grad[0]:
            wt = frozen_transposed(ctx.holder) if (cdt == torch.bfloat16 and ctx.holder is not None) else None
            dx = G.gemm_bf16(dy2, wt) if wt is not None else torch.mm(dy2, w_base.to(cdt))     # frozen base: own GEMM on the transposed copy
            dx.addmm_(da_s, a_w)
            dx = dx.view(ctx.xshape)
        d_a = d_b = None
        if ctx.need_ab:
            g_b = torch.mm(dy2.t(), xa_s)                                    # [N, r]
            g_a = torch.mm(da_s.t(), x2)                                     # [r, K]
            if ctx.sink is not None and dy2.is_cuda and g_a.dtype == torch.bfloat16:
                ctx.sink.append((g_a, ctx.off_a))
                ctx.sink.append((g_b, ctx.off_b))
            else:
                d_a, d_b = g_a.float(), g_b.float()
        return dx, d_a, d_b, None, None, None, None, None, None, None, None


class LoRALinear(nn.Module):
    """y = x W^T + (alpha/r) (x A^T) B^T with W frozen; A,B are the federated parameters."""

    def __init__(self, in_features: int, out_features: int, r: int = 16, alpha: float = 32.0, device=None):
        super().__init__()
        self.base = FrozenLinear(in_features, out_features, device=device)
        self.lora_A = nn.Parameter(torch.empty(r, in_features).normal_(0.0, 1.0 / math.sqrt(in_features)))
        self.lora_B = nn.Parameter(torch.zeros(out_features, r))
        self.scaling = alpha / r
        self.a_bf16: Optional[torch.Tensor] = None      # views into the bf16 shadow buffer (attach_shadow)
        self.b_bf16: Optional[torch.Tensor] = None
        self._sink = None
        self._off_a = self._off_b = 0

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _LoRALinearFn.apply(x, self.lora_A, self.lora_B, self.base.weight_bf16, self.a_bf16, self.b_bf16,
                                   self.scaling, self._sink, self._off_a, self._off_b, self.base)


class FusedLayerNorm(nn.Module):
    def __init__(self, dim: int, eps: float = 1e-12):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.bias = nn.Parameter(torch.zeros(dim))
        self.eps = eps

    def forward(self, x, residual=None):
        """Returns LN(x + residual) (and the summed stream when ``residual`` is given)."""
        y, h = N.layer_norm(x, self.weight, self.bias, self.eps, residual)
        return (y, h) if residual is not None else y


class FusedRMSNorm(nn.Module):
    def __init__(self, dim: int, eps: float = 1e-5):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.eps = eps

    def forward(self, x, residual=None):
        y, h = N.rms_norm(x, self.weight, self.eps, residual)
        return (y, h) if residual is not None else y


def attach_shadow(module: nn.Module, flat_model) -> int:
    """Point every ``ShadowLinear.w_bf16`` / ``ShadowConv2d.w_bf16`` at its slice of the flat bf16 shadow
    buffer (convolutions also get the flat model's gradient sink)."""
    if flat_model.shadow is None:
        return 0
    from .conv import ShadowConv2d

    views = flat_model.shadow_views()
    n = 0
    if hasattr(module, "word_bf16") and "word.weight" in views:      # tied embedding / output projection (models/bert.py)
        module.word_bf16 = views["word.weight"]
        module._sink, module._word_offset = flat_model.grad_sink, flat_model.segment("word.weight").offset
        n += 1
    for name, m in module.named_modules():
        if isinstance(m, LoRALinear):
            ka, kb = (f"{name}.lora_A", f"{name}.lora_B") if name else ("lora_A", "lora_B")
            if ka in views and kb in views:
                m.a_bf16, m.b_bf16 = views[ka], views[kb]
                m._sink = flat_model.grad_sink
                m._off_a, m._off_b = flat_model.segment(ka).offset, flat_model.segment(kb).offset
                n += 1
            continue
        key = f"{name}.weight" if name else "weight"
        if key not in views:
            continue
        if isinstance(m, ShadowLinear):
            m.w_bf16 = views[key]
            m._sink, m._offset = flat_model.grad_sink, flat_model.segment(key).offset
            n += 1
        elif isinstance(m, ShadowConv2d) and (flat_model.segment(key).channels_last or tuple(m.kernel_size) == (1, 1)):
            m.attach(views[key], flat_model.grad_sink, flat_model.segment(key).offset)
            n += 1
    return n


def packed_attention(qkv: torch.Tensor, causal: bool = False, mask=None) -> torch.Tensor:
    """Self-attention on a packed projection output qkv:[B,S,3,H,D] -> [B,S,H,D] (no q/k/v slice copies on the
    tcgen05 path; falls back to :func:`attention` elsewhere)."""
    try:
        from ..ops import attention as A

        q = qkv[:, :, 0]
        if qkv.is_cuda and A.available() and mask is None and A._supported(q, q):
            return A.packed_qkv_attention(qkv, causal)
    except ImportError:
        pass
    return attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], causal, mask)


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, causal: bool, mask=None) -> torch.Tensor:
    """q:[B,S,Hq,D] k,v:[B,S,Hkv,D] -> [B,S,Hq,D].

    K4 (hand-written tcgen05 flash attention) is used when available (ops/attention.py);
    otherwise the library SDPA kernel (same role as cuBLAS for plain GEMMs)."""
    try:
        from ..ops import attention as A

        if q.is_cuda and A.available() and mask is None and A._supported(q, k):
            return A.flash_attention(q, k, v, causal)
    except ImportError:
        pass
    Hq, Hkv = q.shape[2], k.shape[2]
    qt, kt, vt = q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)
    if Hkv != Hq:
        rep = Hq // Hkv
        kt = kt.repeat_interleave(rep, dim=1)
        vt = vt.repeat_interleave(rep, dim=1)
    o = torch.nn.functional.scaled_dot_product_attention(qt, kt, vt, attn_mask=mask, is_causal=causal and mask is None)
    return o.transpose(1, 2)
