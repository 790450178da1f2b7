"""Cross-entropy over bf16 logits without an fp32 copy (csrc/ce.cu): the Llama-3 head produces 1024 x 128256 logits per step; the
composed expression ``cross_entropy(logits.float(), ...)`` moved ~4.4 GB and kept two 525 MB fp32 tensors alive for it.

``fused_cross_entropy(logits [T, V] bf16, labels [T] int64, ignore_index=-100)`` -> mean loss over the counted rows (fp32 scalar).
The backward writes ``dlogits`` IN PLACE over ``logits`` (they must be the freshly computed output of the head GEMM and not be used
again; a second backward through the same graph is not supported).

Reference parity: the reference has no compute; contract = BASELINE.json config 4.
"""
from __future__ import annotations

import torch

from . import count, native, stream_ptr


class _CEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits: torch.Tensor, labels: torch.Tensor, ignore_index: int, n_classes: int):
        T, v_alloc = logits.shape
        V = int(n_classes)
        lse = torch.empty(T, device=logits.device, dtype=torch.float32)
        rows = torch.empty(T, device=logits.device, dtype=torch.float32)
        count(1)
        native().ce_fwd(logits.data_ptr(), labels.data_ptr(), lse.data_ptr(), rows.data_ptr(), T, V, logits.stride(0), int(ignore_index),
                        stream_ptr())
        n = (labels != ignore_index).sum().clamp_(min=1).to(torch.float32)
        ctx.save_for_backward(logits, labels, lse, n)
        ctx.ignore_index, ctx.n_classes = int(ignore_index), V
        return rows.sum() / n

    @staticmethod
    def backward(ctx, g):
        logits, labels, lse, n = ctx.saved_tensors
        T, v_alloc = logits.shape
        scale = (g.to(torch.float32) / n).reshape(1).contiguous()
        count(1)
        native().ce_bwd(logits.data_ptr(), labels.data_ptr(), lse.data_ptr(), scale.data_ptr(), T, ctx.n_classes, v_alloc, logits.stride(0),
                        ctx.ignore_index, stream_ptr())
        return logits, None, None, None    # the logits buffer now holds dlogits (zero in the padding columns)


def fused_cross_entropy(logits: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100, n_classes: int | None = None) -> torch.Tensor:
    """Mean cross-entropy of ``logits`` [T, V] against ``labels`` [T]; rows labelled ``ignore_index`` do not count.
    ``n_classes`` < V: the trailing columns are padding of the head GEMM (vocabulary padded to a tile multiple) -- they take no part
    in the softmax and get a zero gradient."""
    V = logits.shape[1] if n_classes is None else int(n_classes)
    if (logits.is_cuda and logits.dtype == torch.bfloat16 and logits.dim() == 2 and logits.stride(1) == 1 and logits.stride(0) % 8 == 0
            and labels.dtype == torch.int64 and labels.is_contiguous()):
        return _CEFn.apply(logits, labels, ignore_index, V)
    return torch.nn.functional.cross_entropy(logits[:, :V].float(), labels, ignore_index=ignore_index)
