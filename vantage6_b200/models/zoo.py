"""Model zoo: name -> (module factory, forward_loss, synthetic batch factory, training defaults).

Every BASELINE.json config is reachable by name: ``resnet50``, ``bert_base``,
``llama3_8b_lora`` (plus ``*_tiny`` variants that run the same code paths at toy size on CPU and
in smoke tests); the logistic GLM lives in models/glm.py because it is not an ``nn.Module``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Tuple

import torch


@dataclass
class ModelSpec:
    name: str
    build: Callable[[torch.device], torch.nn.Module]
    forward_loss: Callable
    make_batches: Callable[..., List[Tuple[torch.Tensor, torch.Tensor]]]
    optimizer: str = "sgd"
    lr: float = 0.05
    upload: str = "weights_f32"
    shadow_bf16: bool = False
    amp: bool = True
    local_steps: int = 8
    batch: int = 64
    trainer_kwargs: dict = field(default_factory=dict)


def _image_batches(res: int, classes: int):
    def make(n_steps: int, batch: int, seed: int, pin: bool = False):
        g = torch.Generator().manual_seed(seed)
        x = torch.randint(0, 256, (n_steps, batch, 3, res, res), dtype=torch.uint8, generator=g)
        y = torch.randint(0, classes, (n_steps, batch), dtype=torch.int64, generator=g)
        if pin and torch.cuda.is_available():
            x, y = x.pin_memory(), y.pin_memory()
        return [(x[i], y[i]) for i in range(n_steps)]
    return make


def _mlm_batches(vocab: int, seq: int):
    def make(n_steps: int, batch: int, seed: int, pin: bool = False):
        from .bert import synthetic_mlm_batch

        g = torch.Generator().manual_seed(seed)
        out = []
        for _ in range(n_steps):
            ids, labels = synthetic_mlm_batch(vocab, batch, seq, max(1, int(0.15 * batch * seq)), generator=g)
            if pin and torch.cuda.is_available():
                ids, labels = ids.pin_memory(), labels.pin_memory()
            out.append((ids, labels))
        return out
    return make


def _lm_batches(vocab: int, seq: int):
    def make(n_steps: int, batch: int, seed: int, pin: bool = False):
        g = torch.Generator().manual_seed(seed)
        out = []
        for _ in range(n_steps):
            ids = torch.randint(0, vocab, (batch, seq), generator=g)
            if pin and torch.cuda.is_available():
                ids = ids.pin_memory()
            out.append((ids, ids))
        return out
    return make


def _specs() -> Dict[str, ModelSpec]:
    from . import bert, llama, resnet

    cl = torch.channels_last
    return {
        "resnet50": ModelSpec("resnet50", lambda d: resnet.resnet50().to(memory_format=cl), resnet.imagenet_forward_loss,
                              _image_batches(224, 1000), "sgd", 0.05, "weights_f32", True, True, 8, 64,
                              dict(momentum=0.9, weight_decay=1e-4)),
        "resnet_tiny": ModelSpec("resnet_tiny", lambda d: resnet.resnet_tiny(10).to(memory_format=cl),
                                 resnet.imagenet_forward_loss, _image_batches(32, 10), "sgd", 0.05, "weights_f32", True,
                                 True, 2, 8, dict(momentum=0.9)),
        # ResNet at full channel width (64..2048) but one block per stage and 64x64 images: every convolution family of
        # ResNet-50 runs on the tcgen05 kernels (smoke / GPU tests) at a fraction of the cost
        "resnet_mini": ModelSpec("resnet_mini", lambda d: resnet.ResNet((1, 1, 1, 1), 10).to(memory_format=cl),
                                 resnet.imagenet_forward_loss, _image_batches(64, 10), "sgd", 0.05, "weights_f32", True,
                                 True, 2, 16, dict(momentum=0.9)),
        "bert_base": ModelSpec("bert_base", lambda d: bert.bert_base(), bert.bert_forward_loss, _mlm_batches(30522, 128),
                               "adamw", 1e-4, "delta_bf16", True, False, 4, 32, dict(weight_decay=0.01, max_grad_norm=1.0)),
        "bert_tiny": ModelSpec("bert_tiny", lambda d: bert.bert_tiny(), bert.bert_forward_loss, _mlm_batches(512, 32),
                               "adamw", 1e-3, "delta_bf16", True, False, 2, 4, dict(weight_decay=0.01)),
        "bert_small": ModelSpec("bert_small", lambda d: bert.bert_small(), bert.bert_forward_loss, _mlm_batches(1024, 64),
                                "adamw", 1e-3, "delta_bf16", True, False, 2, 8, dict(weight_decay=0.01)),
        "llama3_8b_lora": ModelSpec("llama3_8b_lora", lambda d: llama.llama3_8b_lora(d), llama.llama_forward_loss,
                                    _lm_batches(128256, 1024), "adamw", 2e-4, "delta_bf16", True, False, 2, 1,
                                    dict(weight_decay=0.0, max_grad_norm=1.0, include_buffers=False)),
        "llama_tiny_lora": ModelSpec("llama_tiny_lora", lambda d: llama.llama_tiny_lora(d), llama.llama_forward_loss,
                                     _lm_batches(512, 32), "adamw", 1e-3, "delta_bf16", True, False, 2, 2,
                                     dict(weight_decay=0.0, include_buffers=False)),
    }


def get(name: str) -> ModelSpec:
    specs = _specs()
    if name not in specs:
        raise KeyError(f"unknown model {name!r}; available: {sorted(specs)}")
    return specs[name]


def build_trainer(name: str, *, rank: int, world: int, device, data_plane: str = "auto", server_mode: str = "sharded",
                  server_opt=None, process_group=None, use_cuda_graph=None, fused_local_optimizer: bool = True,
                  **overrides):
    """Construct a :class:`FederatedTrainer` for a zoo model with its training defaults."""
    from ..parallel.trainer import FederatedTrainer
    from .transformer import attach_shadow

    spec = get(name)
    device = torch.device(device)
    model = spec.build(device)
    kw = dict(spec.trainer_kwargs)
    kw.update(overrides)
    if not fused_local_optimizer:
        kw["shadow_bf16"] = False       # torch.optim does not maintain the bf16 shadow: cast per forward
    tr = FederatedTrainer(model, spec.forward_loss, rank=rank, world=world, device=device, optimizer=spec.optimizer,
                          lr=kw.pop("lr", spec.lr), upload=kw.pop("upload", spec.upload),
                          shadow_bf16=kw.pop("shadow_bf16", spec.shadow_bf16),
                          amp_dtype=torch.bfloat16 if spec.amp else None, data_plane=data_plane, server_mode=server_mode,
                          server_opt=server_opt, process_group=process_group, use_cuda_graph=use_cuda_graph,
                          fused_local_optimizer=fused_local_optimizer, **kw)
    attach_shadow(tr.model, tr.fm)
    tr._fresh_model = lambda: spec.build(torch.device("cpu"))      # FederatedTrainer.reset(seed) of a resident trainer
    return tr, spec
