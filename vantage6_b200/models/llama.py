"""Llama-3 8B with LoRA adapters (BASELINE config 4: "Llama-3 8B LoRA federated fine-tune,
8 GPU-nodes, delta = LoRA adapters only").

32 layers, hidden 4096, 32 query / 8 KV heads x 128, SwiGLU FFN 14336, vocab 128256, RoPE theta
5e5, RMSNorm.  The 8 B base weights are frozen bf16 (16 GB, resident in HBM on every node, never
broadcast); the LoRA adapters (r=16 on q,k,v,o: 13.6 M fp32 parameters) are the only trainable,
federated parameters -- 27 MB bf16 per node per round.

Hand-written pieces: base projections forward on the tcgen05 GEMM, RMSNorm fwd+bwd with fused
residual (K5), RoPE in place on q,k (K6), fused flat AdamW (K7), LoRA-delta FedAvg (K2).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn

from ..ops import rope as R
from ..ops.act import swiglu
from .transformer import FrozenLinear, FusedRMSNorm, LoRALinear, attention


@dataclass
class LlamaConfig:
    vocab_size: int = 128256
    hidden: int = 4096
    layers: int = 32
    heads: int = 32
    kv_heads: int = 8
    ffn: int = 14336
    rope_theta: float = 500000.0
    eps: float = 1e-5
    lora_r: int = 16
    lora_alpha: float = 32.0
    max_seq: int = 8192


class LlamaLayer(nn.Module):
    def __init__(self, c: LlamaConfig, device=None):
        super().__init__()
        hd = c.hidden // c.heads
        self.c, self.hd = c, hd
        self.q = LoRALinear(c.hidden, c.heads * hd, c.lora_r, c.lora_alpha, device)
        self.k = LoRALinear(c.hidden, c.kv_heads * hd, c.lora_r, c.lora_alpha, device)
        self.v = LoRALinear(c.hidden, c.kv_heads * hd, c.lora_r, c.lora_alpha, device)
        self.o = LoRALinear(c.heads * hd, c.hidden, c.lora_r, c.lora_alpha, device)
        self.gate = FrozenLinear(c.hidden, c.ffn, device)
        self.up = FrozenLinear(c.hidden, c.ffn, device)
        self.down = FrozenLinear(c.ffn, c.hidden, device)
        self.norm1 = FusedRMSNorm(c.hidden, c.eps)
        self.norm2 = FusedRMSNorm(c.hidden, c.eps)
        for n in (self.norm1, self.norm2):
            n.weight.requires_grad_(False)          # base model is frozen: adapters only

    def forward(self, h, stream, cos, sin):
        """``h`` is the not-yet-added branch output, ``stream`` the residual stream: the add is
        fused into the RMSNorm kernel."""
        c = self.c
        B, S, _ = h.shape
        x, stream = self.norm1(h, residual=stream)
        q = self.q(x).view(B, S, c.heads, self.hd).contiguous()
        k = self.k(x).view(B, S, c.kv_heads, self.hd).contiguous()
        v = self.v(x).view(B, S, c.kv_heads, self.hd)
        q, k = R.apply_rope(q, k, cos, sin)
        a = attention(q, k, v, causal=True).reshape(B, S, -1)
        h = self.o(a)
        x, stream = self.norm2(h, residual=stream)
        h = self.down(swiglu(self.gate(x), self.up(x)))        # one fused kernel each way (ops/act.py)
        return h, stream


class LlamaLoRA(nn.Module):
    def __init__(self, c: LlamaConfig = LlamaConfig(), device=None):
        super().__init__()
        self.cfg = c
        emb = torch.empty(c.vocab_size, c.hidden, dtype=torch.bfloat16, device=device).normal_(0.0, 0.02)
        self.embed_bf16 = emb                                   # frozen, not a parameter
        self.layers = nn.ModuleList(LlamaLayer(c, device) for _ in range(c.layers))
        self.norm = FusedRMSNorm(c.hidden, c.eps)
        self.norm.weight.requires_grad_(False)
        self.lm_head = FrozenLinear(c.hidden, c.vocab_size, device)
        self._rope = None

    def _apply(self, fn, recurse=True):
        super()._apply(fn, recurse)
        self.embed_bf16 = fn(self.embed_bf16)
        self._rope = None
        return self

    def forward(self, input_ids: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        B, S = input_ids.shape
        if self._rope is None or self._rope[0].shape[0] < S or self._rope[0].device != input_ids.device:
            self._rope = R.rope_tables(max(S, 128), self.cfg.hidden // self.cfg.heads, self.cfg.rope_theta,
                                       device=input_ids.device)
        cos, sin = self._rope
        stream = torch.nn.functional.embedding(input_ids, self.embed_bf16)
        h = torch.zeros_like(stream)
        for layer in self.layers:
            h, stream = layer(h, stream, cos, sin)
        x, _ = self.norm(h, residual=stream)
        logits = self.lm_head(x)
        if logits.is_cuda and logits.dtype == torch.bfloat16:
            # next-token targets for every position, the last one ignored: the fused cross-entropy reads the bf16 logits once per
            # direction and writes dlogits in place (ops/ce.py) instead of slicing + an fp32 copy + log_softmax
            from ..ops.ce import fused_cross_entropy

            tgt = torch.cat([labels[:, 1:], torch.full_like(labels[:, :1], -100)], dim=1)
            return fused_cross_entropy(logits.reshape(-1, logits.shape[-1]), tgt.reshape(-1))
        logits = logits.float()
        return torch.nn.functional.cross_entropy(logits[:, :-1].reshape(-1, logits.shape[-1]), labels[:, 1:].reshape(-1))


def llama3_8b_lora(device=None) -> LlamaLoRA:
    return LlamaLoRA(LlamaConfig(), device=device)


def llama_tiny_lora(device=None) -> LlamaLoRA:
    return LlamaLoRA(LlamaConfig(vocab_size=512, hidden=256, layers=2, heads=4, kv_heads=2, ffn=512, lora_r=4,
                                 lora_alpha=8.0), device=device)


def llama_forward_loss(model: nn.Module, input_ids: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    return model(input_ids, labels)
