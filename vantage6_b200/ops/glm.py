"""K8: fused logistic-regression step -- gradient, intercept gradient and loss in one pass over X.

Two kernels: ``csrc/glm_tc.cu`` (both products z = X w and g = X^T r as UMMA GEMVs on one TMA tile of X, the second
reading the tile MN-major; default for bf16, F == 256) and ``csrc/rope_glm.cu::glm_logistic_kernel`` (CUDA cores; any
F % 64 == 0 <= 512, bf16 / fp32).

Returns the *payload vector* ``[g_w (F), g_b, loss_sum, n_rows]`` (padded to a multiple of 4)
that is handed unchanged to the K3 small-message all-reduce: the federated GLM update is
``w <- w - lr * sum_i g_i / sum_i n_i``.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import count, native, stream_ptr

_MAX_PARTS = 148 * 4


def payload_len(F: int) -> int:
    return (F + 3 + 3) // 4 * 4


def logistic_grad(X: torch.Tensor, y: torch.Tensor, w: torch.Tensor, out: Optional[torch.Tensor] = None,
                  scratch: Optional[torch.Tensor] = None) -> torch.Tensor:
    """X:[rows,F] (bf16/fp32), y:[rows] fp32 in {0,1}, w:[F+1] fp32 (last = intercept)."""
    rows, F = X.shape
    if out is None:
        out = torch.zeros(payload_len(F), device=X.device, dtype=torch.float32)
    if X.is_cuda:
        if scratch is None:
            scratch = torch.empty(_MAX_PARTS * (F + 2), device=X.device, dtype=torch.float32)
        count(2)                     # gradient kernel + fold of the per-CTA partials
        if _tensor_core_path(X):     # UMMA formulation (csrc/glm_tc.cu): 99.6 us vs 124.2 us on 1M x 256 bf16
            native().glm_logistic_grad_tc(X.data_ptr(), y.data_ptr(), w.data_ptr(), scratch.data_ptr(), _MAX_PARTS,
                                          out.data_ptr(), rows, F, stream_ptr())
        else:
            native().glm_logistic_grad(X.data_ptr(), y.data_ptr(), w.data_ptr(), scratch.data_ptr(), _MAX_PARTS,
                                       out.data_ptr(), rows, F, X.dtype == torch.bfloat16, stream_ptr())
    else:
        out[: F + 3] = reference_logistic_grad(X, y, w)
    return out


def _tensor_core_path(X: torch.Tensor) -> bool:
    """The tcgen05 kernel handles the benchmark layout (bf16, F == 256, dense rows) and is the default there;
    ``V6B200_GLM=cuda`` forces the CUDA-core kernel, which also covers fp32 inputs and other feature counts."""
    return (os.environ.get("V6B200_GLM", "tc") == "tc" and X.dtype == torch.bfloat16 and X.shape[1] == 256
            and X.is_contiguous() and X.data_ptr() % 16 == 0 and hasattr(native(), "glm_logistic_grad_tc"))


def reference_logistic_grad(X, y, w):
    Xf = X.float()
    F = Xf.shape[1]
    z = Xf @ w[:F].float() + w[F].float()
    r = torch.sigmoid(z) - y.float()
    g = Xf.t() @ r
    loss = torch.nn.functional.binary_cross_entropy_with_logits(z, y.float(), reduction="sum")
    return torch.cat([g, r.sum()[None], loss[None], torch.tensor([float(Xf.shape[0])], device=X.device)])
