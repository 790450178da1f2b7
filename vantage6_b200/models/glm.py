"""Federated logistic-regression GLM (BASELINE config 5: "logistic-regression federated GLM,
tabular 1M x 256 synthetic, 8 GPU-nodes -- exercises the small-message aggregation path").

Each node holds a row shard X_i [n_i, F] (bf16 in HBM, read exactly once per iteration by the
fused K8 kernel) and labels y_i.  One iteration = K8 (gradient + intercept gradient + loss +
count in one 260-float payload) -> K3 (one-shot small all-reduce over NVLink) -> local update
``w <- w - lr * g / n`` (identical on every node, so no broadcast is needed).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ..ops import glm as K8
from ..parallel.fedavg import SmallAggregator


def synthetic_glm_shard(rows: int, features: int = 256, seed: int = 0, device="cpu", dtype=torch.bfloat16,
                        w_true_seed: int = 12345) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(X [rows,F], y [rows] in {0,1}, w_true [F+1]) -- generated directly on ``device``."""
    gw = torch.Generator(device="cpu").manual_seed(w_true_seed)
    w_true = torch.randn(features + 1, generator=gw) * 0.5
    g = torch.Generator(device=device).manual_seed(seed)
    X = torch.randn(rows, features, device=device, generator=g, dtype=torch.float32)
    z = X @ w_true[:features].to(device) + w_true[features].to(device)
    y = (torch.rand(rows, device=device, generator=g) < torch.sigmoid(z)).float()
    return X.to(dtype), y, w_true.to(device)


class FederatedGLM:
    def __init__(self, X: torch.Tensor, y: torch.Tensor, rank: int = 0, world: int = 1, lr: float = 1.0,
                 process_group=None):
        self.X, self.y = X, y
        self.rows, self.F = X.shape
        self.device = X.device
        self.lr = lr
        self.w = torch.zeros(self.F + 1, dtype=torch.float32, device=self.device)
        self.agg = SmallAggregator(K8.payload_len(self.F), rank, world, self.device, process_group=process_group)
        self._scratch = torch.empty(148 * 4 * (self.F + 2), dtype=torch.float32, device=self.device) if X.is_cuda else None
        self.last_loss: Optional[torch.Tensor] = None

    def _fused_ok(self) -> bool:
        """Default (validated on hardware in round 2: 58.4 vs 74.0 us per iteration at 2 GPUs, identical losses;
        ``V6B200_GLM_FUSED=0`` selects the composed path): gradient kernel + ONE kernel that folds its
        partials, all-reduces the payload over NVLink and applies the update (csrc/fedavg.cu::glm_aggregate_update_kernel)."""
        import os

        return (os.environ.get("V6B200_GLM_FUSED", "1") == "1" and self.X.is_cuda and getattr(self.agg, "native", False)
                and K8._tensor_core_path(self.X))

    @torch.no_grad()
    def _step_fused(self) -> torch.Tensor:
        from ..ops import native, stream_ptr

        C, agg = native(), self.agg
        if self.last_loss is None or self.last_loss.dim() != 0 or not self.last_loss.is_cuda:
            self.last_loss = torch.zeros((), dtype=torch.float32, device=self.device)
        from ..ops import count

        count(2)
        nparts = C.glm_logistic_partials_tc(self.X.data_ptr(), self.y.data_ptr(), self.w.data_ptr(), self._scratch.data_ptr(),
                                            148 * 4, self.rows, self.F, stream_ptr())
        agg.epoch += 1
        off = (agg.epoch & 1) * agg.n * 4
        C.glm_aggregate_update(agg._slots.peer(off), agg._pad_buf.peer(), [1.0] * agg.world, agg.out.data_ptr(), agg.n, agg.rank,
                               agg.world, agg.epoch, agg.timeout_cycles, self._scratch.data_ptr(), nparts, self.F,
                               float(self.rows), float(self.lr), self.w.data_ptr(), self.last_loss.data_ptr(), stream_ptr())
        return self.last_loss

    @torch.no_grad()
    def step(self) -> torch.Tensor:
        """One federated gradient step; returns the global mean loss (device scalar)."""
        if self._fused_ok():
            return self._step_fused()
        K8.logistic_grad(self.X, self.y, self.w, out=self.agg.slot(), scratch=self._scratch)
        tot = self.agg.allreduce(1.0, normalize=False)        # sums over nodes: [g_w, g_b, loss, n]
        n = tot[self.F + 2]
        self.w.add_(tot[: self.F + 1] / n, alpha=-self.lr)
        self.last_loss = tot[self.F + 1] / n
        return self.last_loss

    def close(self):
        self.agg.close()
