"""K7: fused flat optimizers (one launch per step) with the federated fusions of SURVEY.md 2.6.

``FlatSGD`` / ``FlatAdamW`` update a contiguous fp32 parameter buffer in one sweep
(csrc/optim.cu).  ``save_ref=True`` on the first local step of a round stores the received
global model; ``publish=...`` on the last local step emits the node's contribution for the
FedAvg reduction (fp32 delta, bf16 delta, or scaled weights) -- the "delta cast/scale" that
the baseline does with separate torch ops.

The pure-PyTorch ``reference_*`` functions are the numerics oracles used by the tests and the
CPU path.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import torch

from . import count, native, stream_ptr

PUBLISH_NONE, PUBLISH_DELTA_F32, PUBLISH_DELTA_BF16, PUBLISH_WEIGHTS_F32 = 0, 1, 2, 3


def _ptr(t: Optional[torch.Tensor]) -> int:
    return 0 if t is None else t.data_ptr()


def _scale_val(cs) -> float:
    """``contrib_scale`` (the node's n_i) is a float or a 1-element device tensor (read by the kernel at run time, so
    a captured CUDA graph multiplies by the CURRENT sample count)."""
    return 1.0 if torch.is_tensor(cs) else float(cs)


def _scale_ptr(cs) -> int:
    return cs.data_ptr() if torch.is_tensor(cs) else 0


@dataclass
class FlatSGD:
    """SGD with momentum / nesterov / weight decay over flat fp32 buffers (torch semantics)."""
    params: torch.Tensor
    lr: float = 0.1
    momentum: float = 0.9
    dampening: float = 0.0
    weight_decay: float = 0.0
    nesterov: bool = False
    steps: int = 0
    buf: torch.Tensor = field(default=None)  # type: ignore[assignment]

    def __post_init__(self):
        assert self.params.dtype == torch.float32 and self.params.is_contiguous()
        if self.buf is None:
            self.buf = torch.zeros_like(self.params)

    def step(self, grads: torch.Tensor, *, w_ref: Optional[torch.Tensor] = None, save_ref: bool = False,
             upload: Optional[torch.Tensor] = None, publish: int = PUBLISH_NONE, contrib_scale: float = 1.0,
             shadow: Optional[torch.Tensor] = None, grad_scale: Optional[torch.Tensor] = None,
             first_momentum_step: Optional[bool] = None) -> None:
        n = self.params.numel()
        first = (self.steps == 0) if first_momentum_step is None else first_momentum_step
        if self.params.is_cuda:
            assert n % 4 == 0, "flat buffers are padded to a multiple of 4 elements"
            count(1)
            native().flat_optim(0, self.params.data_ptr(), grads.data_ptr(), self.buf.data_ptr(), 0, _ptr(w_ref),
                                _ptr(upload), _ptr(shadow), _ptr(grad_scale), n, self.lr, self.momentum,
                                self.dampening, self.weight_decay, 0.0, 0.0, 0.0, 1.0, 1.0, _scale_val(contrib_scale),
                                self.nesterov, save_ref, publish, first, stream_ptr(), 0, _scale_ptr(contrib_scale))
        else:
            reference_sgd_step(self.params, grads, self.buf, self.lr, self.momentum, self.dampening,
                               self.weight_decay, self.nesterov, first, w_ref=w_ref, save_ref=save_ref, upload=upload,
                               publish=publish, contrib_scale=contrib_scale, shadow=shadow, grad_scale=grad_scale)
        self.steps += 1

    def state_dict(self):
        return {"buf": self.buf, "steps": self.steps}

    def load_state_dict(self, sd):
        self.buf.copy_(sd["buf"])
        self.steps = int(sd["steps"])


@dataclass
class FlatAdamW:
    """AdamW (decoupled weight decay) over flat fp32 buffers (torch.optim.AdamW semantics)."""
    params: torch.Tensor
    lr: float = 1e-3
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8
    weight_decay: float = 0.01
    steps: int = 0
    m: torch.Tensor = field(default=None)  # type: ignore[assignment]
    v: torch.Tensor = field(default=None)  # type: ignore[assignment]
    _dev_step: Optional[torch.Tensor] = None
    _dev_bias: Optional[torch.Tensor] = None

    def __post_init__(self):
        assert self.params.dtype == torch.float32 and self.params.is_contiguous()
        if self.m is None:
            self.m = torch.zeros_like(self.params)
        if self.v is None:
            self.v = torch.zeros_like(self.params)

    def step(self, grads: torch.Tensor, *, w_ref: Optional[torch.Tensor] = None, save_ref: bool = False,
             upload: Optional[torch.Tensor] = None, publish: int = PUBLISH_NONE, contrib_scale: float = 1.0,
             shadow: Optional[torch.Tensor] = None, grad_scale: Optional[torch.Tensor] = None,
             device_step: bool = False) -> None:
        self.steps += 1
        bias1 = 1.0 / (1.0 - self.beta1 ** self.steps)
        bias2 = 1.0 / (1.0 - self.beta2 ** self.steps)
        n = self.params.numel()
        if self.params.is_cuda:
            assert n % 4 == 0
            count(1)
            bias_ptr = 0
            if device_step:
                # CUDA-graph safe: the step counter and the bias corrections live on the device
                if self._dev_step is None:
                    self._dev_step = torch.full((1,), self.steps - 1, dtype=torch.int32, device=self.params.device)
                    self._dev_bias = torch.zeros(2, dtype=torch.float32, device=self.params.device)
                count(1)
                native().adam_bias_update(self._dev_step.data_ptr(), self.beta1, self.beta2, self._dev_bias.data_ptr(), stream_ptr())
                bias_ptr = self._dev_bias.data_ptr()
            native().flat_optim(1, self.params.data_ptr(), grads.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                _ptr(w_ref), _ptr(upload), _ptr(shadow), _ptr(grad_scale), n, self.lr, 0.0, 0.0,
                                self.weight_decay, self.beta1, self.beta2, self.eps, bias1, bias2, _scale_val(contrib_scale),
                                False, save_ref, publish, False, stream_ptr(), bias_ptr, _scale_ptr(contrib_scale))
        else:
            reference_adamw_step(self.params, grads, self.m, self.v, self.lr, self.beta1, self.beta2, self.eps,
                                 self.weight_decay, self.steps, w_ref=w_ref, save_ref=save_ref, upload=upload,
                                 publish=publish, contrib_scale=contrib_scale, shadow=shadow, grad_scale=grad_scale)

    def state_dict(self):
        return {"m": self.m, "v": self.v, "steps": self.steps}

    def load_state_dict(self, sd):
        self.m.copy_(sd["m"])
        self.v.copy_(sd["v"])
        self.steps = int(sd["steps"])
        if self._dev_step is not None:
            self._dev_step.fill_(self.steps)


# ----------------------------------------------------------------------------------------------
# reference implementations (plain PyTorch, any device)
# ----------------------------------------------------------------------------------------------
def _publish_ref(w_old, w_new, w_ref, save_ref, upload, publish, contrib_scale, shadow):
    if save_ref:
        w_ref.copy_(w_old)
    if publish in (PUBLISH_DELTA_F32, PUBLISH_DELTA_BF16):
        upload.copy_((contrib_scale * (w_new - w_ref)).to(upload.dtype))
    elif publish == PUBLISH_WEIGHTS_F32:
        upload.copy_(contrib_scale * w_new)
    if shadow is not None:
        shadow.copy_(w_new.to(shadow.dtype))


@torch.no_grad()
def reference_sgd_step(w, g, buf, lr, momentum, dampening, weight_decay, nesterov, first, *, w_ref=None,
                       save_ref=False, upload=None, publish=PUBLISH_NONE, contrib_scale=1.0, shadow=None,
                       grad_scale=None):
    w_old = w.clone()
    g = g * grad_scale if grad_scale is not None else g
    d = g + weight_decay * w
    if momentum != 0.0:
        if first:
            buf.copy_(d)
        else:
            buf.mul_(momentum).add_(d, alpha=1.0 - dampening)
        d = d + momentum * buf if nesterov else buf
    w.add_(d, alpha=-lr)
    _publish_ref(w_old, w, w_ref, save_ref, upload, publish, contrib_scale, shadow)


@torch.no_grad()
def reference_adamw_step(w, g, m, v, lr, beta1, beta2, eps, weight_decay, step, *, w_ref=None, save_ref=False,
                         upload=None, publish=PUBLISH_NONE, contrib_scale=1.0, shadow=None, grad_scale=None):
    w_old = w.clone()
    g = g * grad_scale if grad_scale is not None else g
    w.mul_(1.0 - lr * weight_decay)
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    mh = m / (1.0 - beta1 ** step)
    vh = v / (1.0 - beta2 ** step)
    w.add_(mh / (vh.sqrt() + eps), alpha=-lr)
    _publish_ref(w_old, w, w_ref, save_ref, upload, publish, contrib_scale, shadow)


def clip_grad_coef(grads: torch.Tensor, max_norm: float, scratch: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Global-norm clipping coefficient as a device scalar (consumed via ``grad_scale=``)."""
    if grads.is_cuda:
        scratch = scratch if scratch is not None else torch.empty(2, device=grads.device, dtype=torch.float32)
        count(2)
        native().clip_coef(grads.data_ptr(), grads.numel(), float(max_norm), scratch[0:1].data_ptr(),
                           scratch[1:2].data_ptr(), stream_ptr())
        return scratch[1:2]
    nrm = grads.float().norm()
    return torch.clamp(max_norm / (nrm + 1e-6), max=1.0).reshape(1)


def multi_accumulate(dst: torch.Tensor, items) -> None:
    """dst[off : off + g.numel()] += g for every (g, off) in ``items`` (g: dense bf16, memory order = the
    destination's).  CUDA: one launch per 96 tensors (csrc/optim.cu::multi_accum_kernel)."""
    if not items:
        return
    if dst.is_cuda:
        count(-(-len(items) // 96))
        native().multi_accum_bf16([g.data_ptr() for g, _ in items], [int(o) for _, o in items],
                                  [g.numel() for g, _ in items], dst.data_ptr(), stream_ptr())
    else:
        for g, off in items:
            dst[off: off + g.numel()] += _dense_1d(g).float()


def _dense_1d(g: torch.Tensor) -> torch.Tensor:
    """The tensor's elements in memory order (channels-last 4-D tensors are permuted back first)."""
    if g.dim() == 4 and not g.is_contiguous():
        return g.permute(0, 2, 3, 1).reshape(-1)
    return g.reshape(-1)


def cast_bf16(src: torch.Tensor, dst: torch.Tensor) -> None:
    if src.is_cuda:
        native().cast_bf16(src.data_ptr(), dst.data_ptr(), src.numel(), stream_ptr())
    else:
        dst.copy_(src.to(torch.bfloat16))


def delta_publish(w: torch.Tensor, ref: Optional[torch.Tensor], upload: torch.Tensor, scale: float) -> None:
    if w.is_cuda:
        native().delta_publish(w.data_ptr(), _ptr(ref), upload.data_ptr(), w.numel(), float(scale),
                               upload.dtype == torch.bfloat16, stream_ptr())
    else:
        base = w - ref if ref is not None else w
        upload.copy_((scale * base).to(upload.dtype))
