"""tcgen05/TMEM/TMA GEMM (csrc/gemm.cu) and K1, the broadcast-fused first forward GEMM.

``linear(x, W, bias, act)``        : y = act(x @ W^T + bias), bf16 in / fp32 accumulate in TMEM.
``bcast_linear(x, W_local, W_server_peer, ...)`` : same math, but the weight tiles are pulled by
TMA from the *server GPU's* copy over NVLink (peer-mapped VA) while the MMAs of already-landed
tiles run; the pulled tiles are written back to ``W_local`` so that after the call the node owns
the new global weights of that layer.  This replaces "ncclBroadcast then cuBLAS GEMM".

Backward passes use cuBLAS (torch.matmul) -- plain library GEMMs -- with the activation
derivative recomputed from the saved pre-activation-free formulation (GELU saved input).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import count, native, stream_ptr

ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2


def _check(x: torch.Tensor, W: torch.Tensor):
    assert x.dtype == torch.bfloat16 and W.dtype == torch.bfloat16, "tcgen05 GEMM is bf16 x bf16 -> fp32 acc"
    assert x.stride(-1) == 1 and W.stride(-1) == 1
    K = x.shape[-1]
    assert W.shape[1] == K and K % 8 == 0, "K must be a multiple of 8 (16-byte TMA rows)"


def gemm_bf16(x2: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
              out: Optional[torch.Tensor] = None, variant: Optional[str] = None) -> torch.Tensor:
    """x2:[M,K], W:[N,K] -> [M,N] bf16 (raw op, no autograd)."""
    _check(x2, W)
    M, K = x2.shape
    N = W.shape[0]
    assert N % 8 == 0
    if variant is None and out is None and _igemm_better(M, N, K) and x2.is_contiguous() and W.is_contiguous():
        # few-tile problems: the implicit-GEMM kernel picks its N tile per problem (64 / 128 / 256 columns) and fills the SMs where
        # the fixed 128x256 tile leaves most idle -- 1024x1024x4096: 18.5 vs 27.1 us, 4096x2304x768: 14.7 vs 16.9 us (cuBLAS 14.9),
        # profiles/kernel_bench_linear_fwd_r2.txt
        from . import conv as _conv

        return _conv.linear_fprop(x2, W, bias, act)
    if out is None:
        out = torch.empty(M, N, device=x2.device, dtype=torch.bfloat16)
    count(1)
    C = native()
    # 2-CTA (cta_group::2, 256x256 tiles) when there is enough work to fill 74 SM pairs; the 1-CTA
    # 128x256 kernel otherwise (finer tiles -> better SM fill for small problems).
    two_cta = _two_cta_default(M, N, K) if variant is None else variant == "2cta"
    if two_cta and M >= 256 and N >= 256 and hasattr(C, "gemm2_bf16"):
        C.gemm2_bf16(x2.data_ptr(), W.data_ptr(), out.data_ptr(), 0 if bias is None else bias.data_ptr(),
                     M, N, K, x2.stride(0), W.stride(0), out.stride(0), act, stream_ptr())
    else:
        # (a split-K of the final partial wave -- fp32 reductions into a workspace, last arriver runs the epilogue -- was written and
        # measured: 0.107 vs 0.101 ms on 1024 x 14336 x 4096, and the extra control flow cost the plain path 4 % on the Llama round;
        # removed again, timings in profiles/kernel_bench_gemm_splitk_r2.txt)
        C.gemm_bf16(x2.data_ptr(), W.data_ptr(), out.data_ptr(), 0 if bias is None else bias.data_ptr(),
                    M, N, K, x2.stride(0), W.stride(0), out.stride(0), act, stream_ptr())
    return out


def _igemm_better(M: int, N: int, K: int) -> bool:
    """Measured crossover (scripts/gpu_r2_call16.sh): up to ~400 tiles of 128x256 and K <= 8192 the implicit-GEMM forward kernel is
    the faster of the two tcgen05 GEMMs; large problems stay on the 128x256 / 256x256 kernels of csrc/gemm.cu / gemm2.cu."""
    import os

    if os.environ.get("V6B200_LINEAR_FWD", "auto") == "gemm" or K % 64 != 0 or N % 8 != 0:
        return False
    tiles = ((M + 127) // 128) * ((N + 255) // 256)
    return tiles <= 400 and K <= 8192


def _two_cta_default(M: int, N: int, K: int) -> bool:
    """Measured on B200 (profiles/kernel_bench_r1b.json): the 2-CTA kernel wins when the K loop is long
    enough to amortise the cluster prologue and there are >= ~1 wave of 256x256 tiles (8192^3: 1363 vs 1247
    TFLOP/s, 16384x4096x4096: 1262 vs 1006); the 1-CTA kernel wins or ties on short-K / few-tile problems.
    ``V6B200_GEMM_2CTA=0|1`` forces one kernel."""
    import os

    forced = os.environ.get("V6B200_GEMM_2CTA")
    if forced in ("0", "1"):
        return forced == "1"
    return K >= 2048 and M % 256 == 0 and (M // 256) * ((N + 255) // 256) >= 64


def bias_act_backward(dy2: torch.Tensor, pre: Optional[torch.Tensor], act: int, bias: Optional[torch.Tensor]):
    """Backward of ``y = act(x W^T + b)`` w.r.t. the pre-activation and the bias, ONE kernel (csrc/act.cu):
    returns ``(dpre, db)``; ``dpre`` is ``dy2`` itself when there is no activation; ``db`` is None when it was
    accumulated straight into ``bias.grad`` (flat-buffer models pre-allocate it) or when there is no bias."""
    R, C = dy2.shape
    need_db = bias is not None
    if dy2.is_cuda and dy2.dtype == torch.bfloat16 and C % 64 == 0 and C <= 4096 and (act != ACT_NONE or need_db):
        from . import tree_scratch

        direct = need_db and bias.grad is not None and bias.grad.dtype == torch.float32 and bias.grad.is_contiguous()
        db = None
        if need_db:
            db = bias.grad if direct else torch.empty(C, device=dy2.device, dtype=torch.float32)
        dpre = torch.empty_like(dy2) if act != ACT_NONE else None
        count(1)
        native().bias_act_bwd(dy2.data_ptr(), 0 if pre is None or act == ACT_NONE else pre.data_ptr(),
                              0 if dpre is None else dpre.data_ptr(), 0 if db is None else db.data_ptr(),
                              tree_scratch(dy2.device).data_ptr(), R, C, int(act), bool(direct), stream_ptr())
        return (dpre if dpre is not None else dy2), (None if (direct or not need_db) else db)
    g = dy2
    if act == ACT_GELU:
        p = pre.float()
        g = (dy2.float() * (0.5 * (1.0 + torch.erf(p * 0.7071067811865476))
                            + p * torch.exp(-0.5 * p * p) * 0.3989422804014327)).to(dy2.dtype)
    elif act == ACT_RELU:
        g = dy2 * (pre > 0)
    return g.contiguous(), (g.float().sum(0) if need_db else None)


def bcast_gemm_bf16(x2: torch.Tensor, W_local: torch.Tensor, W_server_ptr: int, ready_flags: torch.Tensor,
                    epoch: int, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """K1. ``W_server_ptr`` is the peer-mapped device address of the server's [N,K] bf16 weights."""
    _check(x2, W_local)
    M, K = x2.shape
    N = W_local.shape[0]
    n_flags = ((N + 255) // 256) * ((K + 63) // 64)
    assert ready_flags.dtype == torch.int32 and ready_flags.numel() >= n_flags
    if out is None:
        out = torch.empty(M, N, device=x2.device, dtype=torch.bfloat16)
    native().bcast_gemm_bf16(x2.data_ptr(), W_local.data_ptr(), W_server_ptr, out.data_ptr(),
                             0 if bias is None else bias.data_ptr(), M, N, K, x2.stride(0), W_local.stride(0),
                             out.stride(0), act, ready_flags.data_ptr(), epoch, stream_ptr())
    return out


def bcast_push_gemm_bf16(x2: torch.Tensor, W_local: torch.Tensor, W_mc_ptr: int, ready_flags: torch.Tensor, flag_peer_ptrs,
                         world: int, is_owner: bool, epoch: int, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
                         out: Optional[torch.Tensor] = None, own_blocks: Optional[tuple] = None, epoch_ptr: int = 0,
                         status_ptr: int = 0) -> torch.Tensor:
    """K1 v3 (push): ``y = x2 @ W^T`` where the new weights are multicast by their owner from inside the GEMM kernel
    (``multimem.st`` through the NVSwitch: one egress copy for any number of nodes) with one ready flag per 256x64 weight
    tile; every rank's tcgen05 main loop consumes tiles from its local copy as their flags arrive -- broadcast and GEMM
    overlap tile by tile, no NCCL call, no separate copy kernel.  ``W_local`` must be this rank's buffer of a symmetric
    allocation bound to the multicast address ``W_mc_ptr``; ``flag_peer_ptrs`` are the peer VAs of every rank's flags.
    ``own_blocks = (lo, hi)``: this rank owns (multicasts) only the 256-row weight blocks [lo, hi) -- a sharded server
    where several ranks hold parts of one layer; ``epoch_ptr``: device word holding the current round (CUDA-graph replays);
    ``status_ptr``: device word set to 1 when a tile never arrives (dead owner) instead of trapping."""
    _check(x2, W_local)
    M, K = x2.shape
    N = W_local.shape[0]
    n_flags = ((N + 255) // 256) * ((K + 63) // 64)
    assert ready_flags.dtype == torch.int32 and ready_flags.numel() >= n_flags and W_mc_ptr
    if out is None:
        out = torch.empty(M, N, device=x2.device, dtype=torch.bfloat16)
    count(1)
    native().bcast_push_gemm_bf16(x2.data_ptr(), W_local.data_ptr(), W_mc_ptr, out.data_ptr(), 0 if bias is None else bias.data_ptr(),
                                  M, N, K, x2.stride(0), W_local.stride(0), out.stride(0), act, ready_flags.data_ptr(),
                                  list(flag_peer_ptrs), world, bool(is_owner), epoch, stream_ptr(),
                                  0 if own_blocks is None else int(own_blocks[0]), -1 if own_blocks is None else int(own_blocks[1]),
                                  int(epoch_ptr), int(status_ptr))
    return out


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, bias, act):
        x2 = x.reshape(-1, x.shape[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        if act == ACT_NONE:
            y = gemm_bf16(x2, W, bias, ACT_NONE)
            ctx.save_for_backward(x2, W, None)
        else:
            pre = gemm_bf16(x2, W, bias, ACT_NONE)          # keep pre-activation for the backward
            y = torch.nn.functional.gelu(pre) if act == ACT_GELU else torch.relu(pre)
            ctx.save_for_backward(x2, W, pre)
        ctx.act = act
        ctx.has_bias = bias is not None
        ctx.xshape = x.shape
        return y.view(*x.shape[:-1], W.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, W, pre = ctx.saved_tensors
        dy2 = dy.reshape(-1, dy.shape[-1])
        if ctx.act == ACT_GELU:
            p = pre.float()
            cdf = 0.5 * (1.0 + torch.erf(p * 0.7071067811865476))
            pdf = torch.exp(-0.5 * p * p) * 0.3989422804014327
            dy2 = (dy2.float() * (cdf + p * pdf)).to(dy.dtype)
        elif ctx.act == ACT_RELU:
            dy2 = dy2 * (pre > 0)
        dx = (dy2 @ W).view(ctx.xshape)
        dW = dy2.t() @ x2
        db = dy2.float().sum(0) if ctx.has_bias else None
        return dx, dW, db, None


class _LinearFusedActFn(torch.autograd.Function):
    """Inference / no-grad fast path: activation fused in the tcgen05 epilogue."""

    @staticmethod
    def forward(ctx, x, W, bias, act):
        x2 = x.reshape(-1, x.shape[-1]).contiguous()
        return gemm_bf16(x2, W, bias, act).view(*x.shape[:-1], W.shape[0])


def linear(x: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE) -> torch.Tensor:
    """y = act(x @ W^T + bias) on the hand-written tcgen05 kernel (CUDA) or torch (CPU)."""
    if x.is_cuda:
        if not torch.is_grad_enabled() or not (x.requires_grad or W.requires_grad):
            return _LinearFusedActFn.apply(x, W, bias, act)
        return _LinearFn.apply(x, W, bias, act)
    return reference_linear(x, W, bias, act)


def reference_linear(x, W, bias=None, act=ACT_NONE):
    y = x.float() @ W.float().t()
    if bias is not None:
        y = y + bias.float()
    if act == ACT_GELU:
        y = torch.nn.functional.gelu(y)
    elif act == ACT_RELU:
        y = torch.relu(y)
    return y.to(x.dtype)
