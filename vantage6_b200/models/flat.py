"""Flat parameter buffers: every model in this framework is "written against flat buffers"
 so that

* the local optimizer is ONE fused kernel over one contiguous fp32 array (ops/optim.py),
* the federated broadcast / reduction moves ONE array that lives in NVLink symmetric memory
  (parallel/fedavg.py) -- the model's ``nn.Parameter``s are *views* into that array, so the new
  global model appears inside the module the moment the aggregation kernel finishes; there is
  no state_dict serialisation, no per-tensor copy.

Layout: ``[trainable params | float buffers (e.g. BatchNorm running stats) | pad]``. FedAvg
averages the whole array; optimizers touch only the trainable prefix.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn


@dataclass
class Segment:
    name: str
    offset: int
    numel: int
    shape: Tuple[int, ...]
    trainable: bool
    channels_last: bool = False      # 4-D tensor stored [O,H,W,I] (cuDNN's NHWC filter layout), exposed as a permuted view


class FlatModel:
    """Re-homes the parameters/buffers of ``module`` into one flat fp32 tensor."""

    ALIGN = 8   # every segment starts on a 32-byte boundary (vector loads in the kernels)

    def __init__(self, module: nn.Module, storage: Optional[torch.Tensor] = None,
                 shadow: Optional[torch.Tensor] = None, include_buffers: bool = True):
        self.module = module
        self.segments: List[Segment] = []
        off = 0
        params = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        # parameters tagged ``_v6_first`` (the weights K1 delivers fused with their first GEMM: parallel/trainer.py
        # ``bcast="fused"``) form one contiguous prefix [0, n_first) of the flat buffers
        params.sort(key=lambda np_: 0 if getattr(np_[1], "_v6_first", False) else 1)
        frozen = [(n, p) for n, p in module.named_parameters() if not p.requires_grad]
        for n, p in params:
            self.segments.append(Segment(n, off, p.numel(), tuple(p.shape), True, _is_channels_last(p)))
            off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.n_trainable = off
        self.n_first = sum((p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN for _, p in params if getattr(p, "_v6_first", False))
        bufs = []
        if include_buffers:
            bufs = [(n, b) for n, b in module.named_buffers() if b.dtype.is_floating_point]
            for n, b in bufs:
                self.segments.append(Segment(n, off, b.numel(), tuple(b.shape), False))
                off += (b.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.n_total = off
        self.frozen = frozen
        device = params[0][1].device if params else torch.device("cpu")
        if storage is None:
            storage = torch.zeros(self.n_total, dtype=torch.float32, device=device)
        assert storage.dtype == torch.float32 and storage.numel() >= self.n_total
        self.flat = storage
        self.grad = torch.zeros(self.n_trainable, dtype=torch.float32, device=storage.device)
        self.shadow = shadow
        self.grad_sink: List[Tuple[torch.Tensor, int]] = []     # (bf16 weight gradient, flat offset) queued by ShadowConv2d
        # move values into the flat storage and alias the module tensors to views
        named = dict(params)
        named_b = dict(bufs)
        with torch.no_grad():
            for seg in self.segments:
                view = self.view_of(self.flat, seg)
                if seg.trainable:
                    p = named[seg.name]
                    view.copy_(p.detach().to(device=self.flat.device, dtype=torch.float32))
                    p.data = view
                    p.grad = self.view_of(self.grad, seg)
                else:
                    b = named_b[seg.name]
                    view.copy_(b.detach().to(device=self.flat.device, dtype=torch.float32))
                    _set_buffer(module, seg.name, view)

    # ------------------------------------------------------------------
    @property
    def params(self) -> torch.Tensor:
        """Trainable prefix of the flat buffer (what the optimizer updates)."""
        return self.flat[: self.n_trainable]

    @staticmethod
    def view_of(buf: torch.Tensor, seg: Segment) -> torch.Tensor:
        """The segment of ``buf`` (any flat buffer with this layout) shaped like the parameter.  Channels-last
        conv filters keep their memory format: no layout-conversion kernel in the step."""
        flat = buf[seg.offset: seg.offset + seg.numel]
        if seg.channels_last:
            o, i, h, w = seg.shape
            return flat.view(o, h, w, i).permute(0, 3, 1, 2)
        return flat.view(seg.shape)

    def zero_grad(self) -> None:
        self.grad.zero_()

    def flush_grad_sink(self) -> int:
        """Add the queued bf16 weight gradients into the flat fp32 gradient buffer with ONE multi-tensor
        kernel (ops/optim.py::multi_accumulate) instead of one cast + one add per layer."""
        n = len(self.grad_sink)
        if n:
            from ..ops import optim as O

            O.multi_accumulate(self.grad, [(g, off) for g, off in self.grad_sink])
            self.grad_sink.clear()
        return n

    def views(self) -> Dict[str, torch.Tensor]:
        return {s.name: self.view_of(self.flat, s) for s in self.segments}

    def shadow_views(self) -> Dict[str, torch.Tensor]:
        assert self.shadow is not None
        return {s.name: self.view_of(self.shadow, s) for s in self.segments}

    def segment(self, name: str) -> Segment:
        for s in self.segments:
            if s.name == name:
                return s
        raise KeyError(name)

    def check_aliasing(self) -> bool:
        """True if every module parameter still aliases the flat storage (debug / tests)."""
        base = self.flat.data_ptr()
        end = base + self.flat.numel() * 4
        return all(base <= p.data_ptr() < end for _, p in self.module.named_parameters() if p.requires_grad)


def _is_channels_last(t: torch.Tensor) -> bool:
    return (t.dim() == 4 and t.is_contiguous(memory_format=torch.channels_last) and not t.is_contiguous())


def flat_size(module: nn.Module, include_buffers: bool = True) -> int:
    """Number of fp32 elements a :class:`FlatModel` of ``module`` needs (before creating storage)."""
    a = FlatModel.ALIGN
    n = sum((p.numel() + a - 1) // a * a for p in module.parameters() if p.requires_grad)
    if include_buffers:
        n += sum((b.numel() + a - 1) // a * a for b in module.buffers() if b.dtype.is_floating_point)
    return n


def _set_buffer(module: nn.Module, dotted: str, value: torch.Tensor) -> None:
    parts = dotted.split(".")
    m = module
    for p in parts[:-1]:
        m = getattr(m, p)
    m._buffers[parts[-1]] = value
