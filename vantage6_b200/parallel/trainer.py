"""FederatedTrainer -- the public training API of one federated node (one GPU, one process).

    trainer = FederatedTrainer(model, forward_loss, rank=r, world=W, device=dev, ...)
    trainer.initialize_global()                    # rank 0's init becomes everyone's model
    loss = trainer.run_round(host_or_device_batches, n_samples)   # E local steps + aggregation

A round = ``len(batches)`` local steps (forward, backward, fused flat optimizer -- captured in
ONE CUDA graph and replayed per step, because a ResNet-50 step is ~500 small launches) followed
by ONE fused aggregation kernel (parallel/fedavg.py).  Input batches may live in pinned host
memory: they are copied host->device on a side stream, double-buffered, overlapping compute.

In vantage6 terms this object is what the ``fedavg`` algorithm's node-side partial function
drives (algorithm/builtin/fedavg.py); task dispatch and result bookkeeping stay in the
server/node control plane.
"""
from __future__ import annotations

import contextlib
from typing import Callable, List, Optional, Sequence, Tuple

import torch
from torch import nn

from ..models import conv as conv_mod
from ..models.flat import FlatModel, flat_size
from ..ops import optim as fused_optim
from .fedavg import FedAvgEngine, ServerOptConfig


class FederatedTrainer:
    def __init__(self, model: nn.Module, forward_loss: Callable[[nn.Module, torch.Tensor, torch.Tensor], torch.Tensor],
                 *, rank: int = 0, world: int = 1, device="cpu", optimizer: str = "sgd", lr: float = 0.1,
                 momentum: float = 0.9, weight_decay: float = 0.0, betas: Tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, server_mode: str = "sharded", server_opt: Optional[ServerOptConfig] = None,
                 upload: str = "weights_f32", data_plane: str = "auto", multicast="auto",
                 use_cuda_graph: Optional[bool] = None, amp_dtype: Optional[torch.dtype] = torch.bfloat16,
                 max_grad_norm: Optional[float] = None, process_group=None, include_buffers: bool = True,
                 shadow_bf16: bool = False, fused_local_optimizer: bool = True, fault_tolerant: bool = False,
                 metrics=None, checkpoint_dir: Optional[str] = None, checkpoint_every: int = 0,
                 timeout_ms: Optional[float] = None, side_wgrad: bool = True, bcast: str = "push"):
        self.rank, self.world = rank, world
        self.device = torch.device(device)
        self.side_wgrad = bool(side_wgrad) and self.device.type == "cuda"
        self.model = model.to(self.device)
        self.forward_loss = forward_loss
        self.amp_dtype = amp_dtype if self.device.type == "cuda" else None
        self.max_grad_norm = max_grad_norm
        n = flat_size(self.model, include_buffers)
        # bcast="fused" (K1): the bf16 weights of the layers tagged ``first_consumer`` are not pushed by the aggregation
        # kernel; they form a prefix of the flat buffers and reach the nodes inside the first forward GEMM that reads them
        # (models/transformer.py::K1_STEP).  Needs the NVLink data plane with multicast, a bf16 shadow, a static set of
        # reducers (not combined with fault_tolerant) and layers whose 256-row blocks line up with the reducer slices.
        from ..models.transformer import ShadowLinear

        self.bcast = bcast
        k1_mods = []
        if bcast == "fused" and world > 1 and shadow_bf16 and self.device.type == "cuda" and not fault_tolerant:
            k1_mods = [m for m in self.model.modules() if isinstance(m, ShadowLinear) and m.first_consumer and m.weight.requires_grad
                       and m.out_features % 256 == 0]
            if len({m.in_features for m in k1_mods}) != 1:
                k1_mods = []
        extra = {}
        if k1_mods:
            for m in k1_mods:
                m.weight._v6_first = True
            n_first = sum((m.weight.numel() + 7) // 8 * 8 for m in k1_mods)
            extra = dict(shard_align=256 * k1_mods[0].in_features, shadow_skip=(0, n_first), shadow_multicast=True)
        self.engine = FedAvgEngine(n, rank, world, self.device, data_plane=data_plane, server_mode=server_mode,
                                   server_opt=server_opt, upload=upload, multicast=multicast,
                                   process_group=process_group, shadow_bf16=shadow_bf16, **extra,
                                   **({"timeout_ms": timeout_ms} if timeout_ms is not None else {}))
        self.fm = FlatModel(self.model, storage=self.engine.w, shadow=self.engine.shadow,
                            include_buffers=include_buffers)
        from ..models.transformer import attach_shadow

        attach_shadow(self.model, self.fm)      # ShadowLinear / ShadowConv2d read the bf16 copy (no-op without a shadow)
        self.k1_layers = 0
        if k1_mods:
            infos = [self.engine.k1_layer(m._offset, m.out_features, m.in_features) for m in k1_mods]
            if all(i is not None for i in infos):
                for m, i in zip(k1_mods, infos):
                    m.k1 = i
                self.k1_layers = len(infos)
            else:       # layout does not allow it: K2 pushes everything, as without bcast="fused"
                self.engine.shadow_skip = (0, 0)
        self.upload_mode = upload
        # delta modes: received global model (trainable prefix saved by the fused optimizer on the
        # first local step; the float-buffer tail, e.g. BatchNorm statistics, saved per round below)
        self.w_ref = torch.zeros(self.fm.n_total, dtype=torch.float32, device=self.device) \
            if upload != "weights_f32" else None
        self.fused_local_optimizer = fused_local_optimizer
        if fused_local_optimizer:
            if optimizer == "sgd":
                self.opt = fused_optim.FlatSGD(self.fm.params, lr=lr, momentum=momentum, weight_decay=weight_decay)
            elif optimizer == "adamw":
                self.opt = fused_optim.FlatAdamW(self.fm.params, lr=lr, beta1=betas[0], beta2=betas[1], eps=eps,
                                                 weight_decay=weight_decay)
            else:
                raise ValueError(optimizer)
            self.torch_opt = None
        else:       # baseline: stock torch.optim over the per-tensor views
            params = [p for p in self.model.parameters() if p.requires_grad]
            self.torch_opt = (torch.optim.SGD(params, lr=lr, momentum=momentum, weight_decay=weight_decay)
                              if optimizer == "sgd" else
                              torch.optim.AdamW(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
            self.opt = None
        self.use_graph = (self.device.type == "cuda") if use_cuda_graph is None else use_cuda_graph
        self._graphs: dict = {}
        self._static_x: Optional[torch.Tensor] = None
        self._static_y: Optional[torch.Tensor] = None
        self.loss_sum = torch.zeros((), dtype=torch.float32, device=self.device)
        self._clip_scratch = torch.zeros(2, dtype=torch.float32, device=self.device) if max_grad_norm else None
        # delta uploads are published pre-multiplied by this node's sample count n_i (a device scalar, so the captured
        # last-step graph picks up the current value): the reducers then only sum -- in the switch when NVLS is there --
        # and divide by sum_i n_i, whatever the n_i are
        self.n_i = torch.ones(1, dtype=torch.float32, device=self.device)
        self._n_host = 1.0
        self.fault_tolerant = fault_tolerant
        self.metrics = metrics
        self.checkpoint_dir, self.checkpoint_every = checkpoint_dir, int(checkpoint_every)
        self.dead: List[int] = []
        self.copy_stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None
        self._staging: List[Tuple[torch.Tensor, torch.Tensor]] = []
        self.native_launches = 0            # launches of OUR kernels so far (graph replays included)
        self._graph_launches: dict = {}
        self.rounds = 0

    # ------------------------------------------------------------------ local step
    def _step_body(self, x: torch.Tensor, y: torch.Tensor, variant: str) -> None:
        self.fm.zero_grad()
        ctx = torch.autocast("cuda", dtype=self.amp_dtype) if self.amp_dtype is not None else contextlib.nullcontext()
        from ..models import transformer as tfm

        tfm.K1_STEP["on"] = bool(self.k1_layers) and variant.endswith("+k1")
        try:
            with ctx:
                loss = self.forward_loss(self.model, x, y)
        finally:
            tfm.K1_STEP["on"] = False
        variant = variant.replace("+k1", "")
        if self.side_wgrad:
            conv_mod.side_wgrad(True)      # filter gradients on a second stream (models/conv.py::_SideWgrad)
        try:
            loss.backward()
        finally:
            if self.side_wgrad:
                conv_mod.join_side_wgrad()
                conv_mod.side_wgrad(False)
        self.fm.flush_grad_sink()          # bf16 conv weight gradients -> flat fp32 grads, one multi-tensor kernel
        self.loss_sum += loss.detach().float()
        if self.torch_opt is not None:
            # baseline arm: stock optimizer; delta upload modes save the round's reference / publish the delta with
            # separate passes (the fused optimizer does both inside its one sweep)
            nt = self.fm.n_trainable
            delta_mode = self.upload_mode != "weights_f32"
            if delta_mode and variant in ("first", "only"):
                self.w_ref[:nt].copy_(self.fm.params)
            if self.max_grad_norm:
                torch.nn.utils.clip_grad_norm_([p for p in self.model.parameters() if p.requires_grad], self.max_grad_norm)
            self.torch_opt.step()
            if delta_mode and variant in ("last", "only"):
                fused_optim.delta_publish(self.fm.params, self.w_ref[:nt], self.engine.upload[:nt], self._n_host)
            return
        gs = None
        if self.max_grad_norm:
            gs = fused_optim.clip_grad_coef(self.fm.grad, self.max_grad_norm, self._clip_scratch)
        kw = {}
        if self.upload_mode != "weights_f32":
            publish = fused_optim.PUBLISH_DELTA_F32 if self.upload_mode == "delta_f32" else fused_optim.PUBLISH_DELTA_BF16
            first, last = variant in ("first", "only"), variant in ("last", "only")
            kw = dict(w_ref=self.w_ref[: self.fm.n_trainable], save_ref=first, upload=self.engine.upload[: self.fm.n_trainable] if last else None,
                      publish=publish if last else fused_optim.PUBLISH_NONE, contrib_scale=self.n_i)
        shadow = self.engine.shadow[: self.fm.n_trainable] if self.engine.shadow is not None else None
        if isinstance(self.opt, fused_optim.FlatSGD):
            self.opt.step(self.fm.grad, grad_scale=gs, shadow=shadow, first_momentum_step=False if self.use_graph else None, **kw)
        else:
            # device_step: Adam's step counter / bias corrections live on the GPU so the captured
            # graph advances them on every replay
            self.opt.step(self.fm.grad, grad_scale=gs, shadow=shadow, device_step=bool(self.use_graph), **kw)

    def _variant(self, i: int, n: int) -> str:
        k1 = "+k1" if (self.k1_layers and i == 0) else ""      # first step of the round: K1 delivers the tagged weights
        if self.upload_mode == "weights_f32":
            return "mid" + k1
        if n == 1:
            return "only" + k1
        return ("first" if i == 0 else ("last" if i == n - 1 else "mid")) + k1

    def _capture(self, variant: str, x: torch.Tensor, y: torch.Tensor) -> torch.cuda.CUDAGraph:
        if self._static_x is None:
            self._static_x, self._static_y = x.clone(), y.clone()
        # snapshot mutable state, warm up on a side stream, restore
        snap = (self.engine.w.clone(), self.opt.state_dict() if self.opt else None, self.loss_sum.clone())
        snap_opt = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in snap[1].items()} if snap[1] else None
        # stock-optimizer arm (comparator "stock_graph"): its per-parameter state must be restored IN PLACE, the captured
        # graph keeps pointing at these buffers
        snap_torch = None
        if self.torch_opt is not None:
            snap_torch = {id(p): {k: v.clone() for k, v in st.items() if torch.is_tensor(v)} for p, st in self.torch_opt.state.items()}
        s = torch.cuda.Stream(self.device)
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            for _ in range(2):
                self._step_body(self._static_x, self._static_y, variant)
        torch.cuda.current_stream(self.device).wait_stream(s)
        g = torch.cuda.CUDAGraph()
        pool = next(iter(self._graphs.values())).pool() if self._graphs else None
        from ..ops import LAUNCHES

        before = LAUNCHES[0]
        with torch.cuda.graph(g, pool=pool):
            self._step_body(self._static_x, self._static_y, variant)
        self._graph_launches[variant] = LAUNCHES[0] - before      # our kernels inside one replay
        self.engine.w.copy_(snap[0])
        if variant.endswith("+k1"):
            # the warm-up runs multicast the owners' (locally advanced) shadow tiles into every rank: restore only after
            # EVERY rank is through its capture, and let nobody start the round before everybody has restored
            torch.cuda.synchronize(self.device)
            self.engine.heap.host_barrier()
        if self.engine.shadow is not None:          # the warm-up / capture steps advanced the bf16 copy as well
            self.engine.shadow.copy_(self.engine.w.to(torch.bfloat16))
        if variant.endswith("+k1"):
            torch.cuda.synchronize(self.device)
            self.engine.heap.host_barrier()
        if snap_opt is not None:
            self.opt.load_state_dict(snap_opt)
        if self.torch_opt is not None:
            for p, st in self.torch_opt.state.items():
                saved = snap_torch.get(id(p), {})
                for k, v in st.items():
                    if torch.is_tensor(v):
                        if k in saved:
                            v.copy_(saved[k])
                        else:
                            v.zero_()       # state created by the warm-up steps (e.g. momentum_buffer): momentum*0 + g == torch's first step
        self.loss_sum.copy_(snap[2])
        return g

    def local_step(self, x: torch.Tensor, y: torch.Tensor, i: int = 0, n: int = 1) -> None:
        """One local optimisation step on device-resident (x, y)."""
        variant = self._variant(i, n)
        if self.use_graph:
            if variant not in self._graphs:
                self._graphs[variant] = self._capture(variant, x, y)
            self._static_x.copy_(x, non_blocking=True)
            self._static_y.copy_(y, non_blocking=True)
            self._graphs[variant].replay()
            self.native_launches += self._graph_launches.get(variant, 0)
            if self.opt is not None:
                self.opt.steps += 1
        else:
            from ..ops import LAUNCHES

            before = LAUNCHES[0]
            self._step_body(x, y, variant)
            self.native_launches += LAUNCHES[0] - before

    # ------------------------------------------------------------------ rounds
    def initialize_global(self) -> None:
        self.engine.initialize_global()

    def remember_init(self, seed: int = 0) -> None:
        """Snapshot the freshly initialised model (a resident trainer restarts later tasks from it, see :meth:`reset`)."""
        self._init_state = (int(seed), self.engine.w.detach().clone())

    @torch.no_grad()
    def reset(self, seed: int = 0) -> None:
        """Start a NEW federated run on this resident trainer (node/gpu_worker.py keeps trainers across tasks): model
        back to its initial weights, local and server optimizer state cleared, round counters zeroed.  The symmetric
        heap, the flat views, the bf16 shadow and the captured CUDA graphs are reused as they are."""
        init = getattr(self, "_init_state", None)
        fresh_fn = getattr(self, "_fresh_model", None)
        if init is not None and (init[0] == int(seed) or fresh_fn is None):
            self.engine.w.copy_(init[1])                   # same seed (or no builder): the remembered initialisation
        elif fresh_fn is not None:
            torch.manual_seed(int(seed))                   # a different seed: draw a new initialisation on the host
            fresh = fresh_fn()
            src = dict(fresh.named_parameters())
            src.update(dict(fresh.named_buffers()))
            for name, view in self.fm.views().items():
                if name in src:
                    view.copy_(src[name].detach().to(device=view.device, dtype=torch.float32))
            self._init_state = (int(seed), self.engine.w.detach().clone())
        if self.engine.shadow is not None:
            self.engine.shadow.copy_(self.engine.w.to(torch.bfloat16))
        if self.opt is not None:
            for v in self.opt.state_dict().values():
                if torch.is_tensor(v):
                    v.zero_()
            self.opt.steps = 0
            if getattr(self.opt, "_dev_step", None) is not None:
                self.opt._dev_step.fill_(-1)
        if self.torch_opt is not None:
            self.torch_opt.state.clear()
        eng = self.engine
        eng.w_global.zero_(); eng.opt_m.zero_(); eng.opt_v.zero_()
        eng.server_step = 0
        if self.w_ref is not None:
            self.w_ref.zero_()
        self.loss_sum.zero_()
        self.rounds = 0

    def _stage(self, k: int, x: torch.Tensor, y: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        while len(self._staging) <= k:
            self._staging.append((torch.empty_like(x, device=self.device), torch.empty_like(y, device=self.device)))
        return self._staging[k]

    def run_round(self, batches: Sequence[Tuple[torch.Tensor, torch.Tensor]], n_samples: Optional[float] = None,
                  weights: Optional[Sequence[float]] = None) -> torch.Tensor:
        """E = len(batches) local steps, then the fused FedAvg aggregation.

        ``batches`` may be pinned-host tensors (copied H2D on a side stream, double-buffered) or
        device tensors.  Returns the mean local loss of the round as a device scalar; calling
        ``.item()`` on it is the round's only device->host read.
        """
        n = len(batches)
        if n_samples is None:
            n_samples = float(sum(b[0].shape[0] for b in batches))
        if self.fault_tolerant and self.rounds > 0:
            self.recover_if_failed()                    # the previous round's status word (GPU idle here: the caller read the loss)
        if float(n_samples) != self._n_host:
            self._n_host = float(n_samples)
            self.n_i.fill_(self._n_host)                # read by the (captured) publish of the last local step
        self.loss_sum.zero_()
        nt, na = self.fm.n_trainable, self.fm.n_total
        has_tail = self.upload_mode != "weights_f32" and na > nt
        if has_tail:        # buffers (BN running stats) are federated too: remember their incoming value
            self.w_ref[nt:na].copy_(self.fm.flat[nt:na])
        cur =torch.cuda.current_stream(self.device) if self.device.type == "cuda" else None
        on_host = n > 0 and batches[0][0].device.type == "cpu" and self.device.type == "cuda"
        events: List[Optional[torch.cuda.Event]] = [None, None]
        done: List[Optional[torch.cuda.Event]] = [None, None]

        def prefetch(i: int):
            k = i & 1
            xs, ys = self._stage(k, *batches[i])
            with torch.cuda.stream(self.copy_stream):
                if done[k] is not None:
                    self.copy_stream.wait_event(done[k])      # staging slot no longer read by compute
                xs.copy_(batches[i][0], non_blocking=True)
                ys.copy_(batches[i][1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
            events[k] = ev

        if on_host:
            prefetch(0)
        for i in range(n):
            if on_host:
                if i + 1 < n:
                    prefetch(i + 1)
                k = i & 1
                cur.wait_event(events[k])
                x, y = self._staging[k]
            else:
                x, y = batches[i]
            self.local_step(x, y, i, n)
            if on_host:
                d = torch.cuda.Event()
                d.record(cur)
                done[i & 1] = d
        if has_tail:        # publish the buffer deltas next to the parameter deltas of the last step
            fused_optim.delta_publish(self.fm.flat[nt:na], self.w_ref[nt:na], self.engine.upload[nt:na], self._n_host)
        # `n_samples` is THIS node's sample count: the engine moves it to the reducers with the contribution (a node
        # never needs to know the other nodes' counts).  An explicit `weights` vector (identical on all ranks) overrides it.
        self._last_agg = (weights, float(n_samples))
        self._aggregate()
        if self.engine.data_plane == "native":
            self.native_launches += 1 + (1 if has_tail else 0)
        self.rounds += 1
        loss = self.loss_sum / max(n, 1)
        if self.metrics is not None:
            self.metrics.log_round(self.rounds, rank=self.rank, world=self.world, local_steps=n, n_samples=float(n_samples),
                                   dead=list(self.dead), upload=self.upload_mode,
                                   nvlink_bytes=self.engine.nvlink_bytes_per_round() if hasattr(self.engine, "nvlink_bytes_per_round") else 0)
        if self.checkpoint_dir and self.checkpoint_every > 0 and self.rounds % self.checkpoint_every == 0:
            self.save_checkpoint()
        return loss

    def _aggregate(self) -> None:
        weights, n_samples = self._last_agg
        delta = self.upload_mode != "weights_f32"
        if weights is not None:
            # explicit vector: in delta modes the contributions are already multiplied by this rank's n_i = n_samples
            if delta:
                self.engine.aggregate([w if w > 0 else 0.0 for w in weights], prescaled=True)
            else:
                self.engine.aggregate(weights)
        else:
            self.engine.aggregate(n_samples, prescaled=delta)

    # ------------------------------------------------------------------ fault handling
    def recover_if_failed(self) -> List[int]:
        """Turn a timed-out aggregation into partial participation: read the status word of the last round; if a
        contributor never arrived, every survivor marks the same ranks dead (the missing set travels between the
        reducers inside the kernel), the slices are re-partitioned, and the aggregation of that round is re-run over
        the survivors -- their contributions are still in place because a reducer that misses a contributor pushes
        nothing.  Returns the ranks newly marked dead."""
        if self.engine.poll_status() == 0:
            return []
        mask = self.engine.missing_mask()
        newly = [r for r in range(self.world) if (mask >> r) & 1 and r != self.rank and r not in self.dead]
        self.engine.clear_status()
        for r in newly:
            self.engine.mark_dead(r, master_hint=self.w_ref)
            self.dead.append(r)
        if newly and getattr(self, "_last_agg", None) is not None:
            weights, n_samples = self._last_agg
            if weights is not None:
                self._last_agg = ([0.0 if i in self.dead else w for i, w in enumerate(weights)], n_samples)
            self._aggregate()
            if self.device.type == "cuda":
                torch.cuda.synchronize(self.device)
            if self.engine.poll_status() != 0:
                raise RuntimeError(f"rank {self.rank}: aggregation still failing after excluding ranks {self.dead}")
        if self.metrics is not None and newly:
            self.metrics.log_round(self.rounds, rank=self.rank, event="recovered", dead=list(self.dead))
        return newly

    # ------------------------------------------------------------------ checkpoint / resume
    def save_checkpoint(self, directory: Optional[str] = None) -> str:
        from ..utils.checkpoint import save_checkpoint

        return save_checkpoint(directory or self.checkpoint_dir, self, self.rounds)

    def load_checkpoint(self, directory: Optional[str] = None) -> int:
        from ..utils.checkpoint import load_checkpoint

        self.rounds = int(load_checkpoint(directory or self.checkpoint_dir, self)["round"])
        return self.rounds

    def launches_per_round(self, n_steps: int) -> int:
        """Number of OUR kernels launched in one round (bench.py ``gpu_launches``): per local step the
        fused flat optimizer (+2 for grad clipping); per round the fused aggregation kernel."""
        per_step = (1 if self.opt is not None else 0) + (2 if (self.max_grad_norm and self.opt is not None) else 0)
        return n_steps * per_step + (1 if self.engine.data_plane == "native" else 0)

    def close(self) -> None:
        self._graphs.clear()
        self.engine.close()
