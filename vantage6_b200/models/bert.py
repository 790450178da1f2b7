"""BERT-base (BASELINE config 3: "BERT-base FedAvg bf16, 8 GPU-nodes, 4 local steps/round").

12 layers, hidden 768, 12 heads x 64, FFN 3072, vocab 30522, post-LayerNorm; masked-LM head with
the decoder tied to the word embeddings.  110 M parameters = 220 MB of bf16 deltas per node per
round (BASELINE.md roofline table).

Hand-written pieces on the hot path: QKV / output / FFN GEMMs forward on the tcgen05 kernel
(bias + GELU epilogue), LayerNorm fwd+bwd with fused residual add (K5), fused flat AdamW with
the delta publish (K7), bf16-delta FedAvg reduction + broadcast (K2).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
from torch import nn

from ..ops import gemm as G
from .transformer import FusedLayerNorm, ShadowLinear, packed_attention


@dataclass
class BertConfig:
    vocab_size: int = 30522
    hidden: int = 768
    layers: int = 12
    heads: int = 12
    ffn: int = 3072
    max_pos: int = 512
    type_vocab: int = 2
    eps: float = 1e-12


class BertLayer(nn.Module):
    def __init__(self, c: BertConfig):
        super().__init__()
        self.heads, self.hd = c.heads, c.hidden // c.heads
        self.qkv = ShadowLinear(c.hidden, 3 * c.hidden, first_consumer=True)      # first GEMM of the layer: K1 candidate
        self.out = ShadowLinear(c.hidden, c.hidden)
        self.ln1 = FusedLayerNorm(c.hidden, c.eps)
        self.ffn1 = ShadowLinear(c.hidden, c.ffn, act=G.ACT_GELU)
        self.ffn2 = ShadowLinear(c.ffn, c.hidden)
        self.ln2 = FusedLayerNorm(c.hidden, c.eps)

    def forward(self, x, mask=None):
        B, S, H = x.shape
        qkv = self.qkv(x).view(B, S, 3, self.heads, self.hd)
        a = packed_attention(qkv, mask=mask).reshape(B, S, H)
        x, _ = self.ln1(self.out(a), residual=x)           # LN(x + attn) with the add fused in the kernel
        x, _ = self.ln2(self.ffn2(self.ffn1(x)), residual=x)
        return x


class BertForMaskedLM(nn.Module):
    def __init__(self, c: BertConfig = BertConfig()):
        super().__init__()
        self.cfg = c
        # the vocabulary is padded to a multiple of 64 rows (30522 -> 30528; the extra rows are never looked up and never enter the
        # softmax): the tied output projection then runs on the tcgen05 kernels (forward, dX, dW) instead of the unaligned-N library
        # fallback (an sm_80 mma.sync kernel at ~180 TFLOP/s, profiles/launches_bert_base_r2_final.txt)
        self.vocab_padded = (c.vocab_size + 63) // 64 * 64
        self.word = nn.Embedding(self.vocab_padded, c.hidden)
        self.word_bf16 = None                    # bf16 shadow of the embedding table (set by attach_shadow)
        self._sink, self._word_offset = None, 0
        self.pos = nn.Embedding(c.max_pos, c.hidden)
        self.tok_type = nn.Embedding(c.type_vocab, c.hidden)
        for e in (self.word, self.pos, self.tok_type):
            nn.init.normal_(e.weight, 0.0, 0.02)
        self.emb_ln = FusedLayerNorm(c.hidden, c.eps)
        self.layers = nn.ModuleList(BertLayer(c) for _ in range(c.layers))
        self.head_dense = ShadowLinear(c.hidden, c.hidden, act=G.ACT_GELU)
        self.head_ln = FusedLayerNorm(c.hidden, c.eps)
        self.head_bias = nn.Parameter(torch.zeros(self.vocab_padded))

    def forward(self, input_ids: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        B, S = input_ids.shape
        pos = torch.arange(S, device=input_ids.device)
        h = self.word(input_ids) + self.pos(pos)[None] + self.tok_type.weight[0][None, None]
        x = self.emb_ln(h.to(torch.bfloat16))
        for layer in self.layers:
            x = layer(x)
        # MLM head only on the masked positions. ``labels`` is [n_masked, 2] = (flat position, token)
        # so every shape is static (CUDA-graph friendly, no nonzero() sync).
        xs = x.reshape(B * S, -1).index_select(0, labels[:, 0])
        t = self.head_ln(self.head_dense(xs))
        V = self.cfg.vocab_size
        if t.is_cuda and t.dtype == torch.bfloat16:
            # tied output projection on the tcgen05 GEMMs (bias in the epilogue), cross-entropy on csrc/ce.cu over the bf16 logits
            from ..ops.ce import fused_cross_entropy
            from .transformer import _ShadowLinearFn

            logits = _ShadowLinearFn.apply(t, self.word.weight, self.head_bias, self.word_bf16, G.ACT_NONE, self._sink, self._word_offset)
            return fused_cross_entropy(logits, labels[:, 1].contiguous(), n_classes=V)
        logits = torch.nn.functional.linear(t, self.word.weight[:V].to(t.dtype)).float() + self.head_bias[:V]
        return torch.nn.functional.cross_entropy(logits, labels[:, 1])


def bert_base() -> BertForMaskedLM:
    return BertForMaskedLM(BertConfig())


def bert_tiny() -> BertForMaskedLM:
    return BertForMaskedLM(BertConfig(vocab_size=512, hidden=64, layers=2, heads=2, ffn=256, max_pos=64))


def bert_small() -> BertForMaskedLM:
    """4 layers of width 256 (QKV = 768 rows = three 256-row blocks): the smallest shape the K1 path takes; multi-GPU checks."""
    return BertForMaskedLM(BertConfig(vocab_size=1024, hidden=256, layers=4, heads=4, ffn=1024, max_pos=128))


def bert_forward_loss(model: nn.Module, input_ids: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    return model(input_ids, labels)


def synthetic_mlm_batch(vocab: int, batch: int, seq: int, n_masked: int, generator=None, device="cpu"):
    """Random token ids; exactly ``n_masked`` positions per batch are masked.  Returns
    ``(input_ids [B,S], labels [n_masked, 2])`` with labels = (flat position, original token)."""
    ids = torch.randint(5, vocab, (batch, seq), generator=generator)
    perm = torch.randperm(batch * seq, generator=generator)[:n_masked].sort().values
    labels = torch.stack([perm, ids.reshape(-1)[perm]], dim=1).to(torch.int64)
    ids.reshape(-1)[perm] = 4          # [MASK]
    return ids.to(device), labels.to(device)
