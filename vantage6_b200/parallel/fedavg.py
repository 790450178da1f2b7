"""FedAvg aggregation engine -- the server's per-round loop on the GPUs.

Data planes
-----------
``native``      ONE fused kernel per round per rank (csrc/fedavg.cu, K2): flag handshake ->
                weighted reduction over NVLink (P2P loads or in-switch ``multimem.ld_reduce``)
                -> server optimizer (FedAvg / FedAvgM / FedAdam) on the fp32 master ->
                broadcast of the new global model fused in the epilogue (P2P stores or
                ``multimem.st``).  No NCCL call.
``collective``  the baseline / CPU path: ``torch.distributed.reduce`` of pre-scaled
                contributions + torch optimizer math + ``torch.distributed.broadcast``
                (NCCL on GPUs, gloo on CPU).  "A path that only calls NCCL for the named ops is
                the baseline, not the product" (BASELINE.md).

Server placement
----------------
``central``  the whole aggregation runs on rank 0 / GPU 0 (the vantage6 central server);
``sharded``  every rank owns 1/world of the global model and of the server-optimizer state
             (the same kernel with a different slice) -- NVLink traffic is balanced over all
             18x8 links instead of funnelled through GPU 0.

Semantics kept from vantage6: node-local data never leaves the node,
only model parameters / deltas cross; weights ``n_i`` are the node sample counts; partial
participation renormalises over the reporters (weight 0 == not reporting).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import torch

from .symm import PAD_WORDS, SymmetricHeap

SERVER_OPTS = {"fedavg": 0, "fedavgm": 1, "fedadam": 2}
UPLOAD_MODES = ("weights_f32", "delta_f32", "delta_bf16")


def device_clock_khz(device) -> int:
    """SM clock the in-kernel timeouts (clock64 cycles) are sized against: the device's maximum, queried -- not assumed."""
    props = torch.cuda.get_device_properties(device)
    khz = getattr(props, "clock_rate", 0)
    if not khz:
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(torch.device(device).index or 0)
            khz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM) * 1000
        except Exception:  # noqa: BLE001
            khz = 2_000_000
    return int(khz)


@dataclass
class ServerOptConfig:
    name: str = "fedavg"
    lr: float = 1.0
    beta1: float = 0.9
    beta2: float = 0.99
    eps: float = 1e-3


def slice_bounds(n: int, n_reducers: int, pos: Optional[int], align: int) -> Tuple[int, int, int]:
    """Slice [lo, hi) of the ``pos``-th of ``n_reducers`` reducers over ``n`` elements (``pos=None``: not a reducer, empty
    slice) and the chunk length: equal chunks rounded up to ``align`` elements, the last ones possibly short or empty.  Over
    all positions the slices tile [0, n) exactly."""
    chunk = (n + n_reducers - 1) // n_reducers
    chunk = (chunk + align - 1) // align * align
    if pos is None:
        return 0, 0, chunk
    lo = min(n, pos * chunk)
    return lo, min(n, lo + chunk), chunk


def k1_own_blocks(lo: int, hi: int, chunk: int, n_reducers: int, offset: int, n_out: int, k_in: int) -> Optional[Tuple[int, int]]:
    """Which 256-row blocks of the [n_out, k_in] weight at flat ``offset`` lie in the reducer slice [lo, hi): a half-open
    block range, (0, 0) when none.  ``None`` -- on every rank alike, it only depends on the shared layout -- when some
    reducer boundary cuts a block (then no single owner could multicast it)."""
    blk, size = 256 * k_in, n_out * k_in
    for pos in range(1, n_reducers):
        b = pos * chunk - offset
        if 0 < b < size and b % blk:
            return None
    a, z = max(lo, offset) - offset, min(hi, offset + size) - offset
    return (a // blk, z // blk) if z > a else (0, 0)


class FedAvgEngine:
    """Owns the symmetric buffers of one rank and runs aggregation rounds.

    Attributes
    ----------
    w : torch.Tensor
        fp32 parameter buffer of this node (length ``n``) living in symmetric memory.  The local
        model's parameters are views into it; after :meth:`aggregate` it holds the new global model.
    upload : torch.Tensor
        contribution buffer (aliases ``w`` in ``weights_f32`` mode).
    """

    def __init__(self, n_params: int, rank: int = 0, world: int = 1, device="cpu", *, data_plane: str = "auto",
                 server_mode: str = "sharded", server_opt: ServerOptConfig | None = None, upload: str = "weights_f32",
                 shadow_bf16: bool = False, multicast: bool | str = "auto", timeout_ms: float = 20000.0,
                 process_group=None, shard_align: int = 8, shadow_skip: Tuple[int, int] = (0, 0), shadow_multicast: bool = False):
        """``shard_align``: reducer slice boundaries are multiples of this many elements; ``shadow_skip`` = (lo, hi): the bf16
        shadow elements K2 leaves on their owner (K1 -- :meth:`k1_layer` -- delivers them inside the first GEMM that reads
        them); ``shadow_multicast``: bind the shadow buffer to a multicast address even when K2 itself runs on P2P."""
        assert upload in UPLOAD_MODES and server_mode in ("central", "sharded")
        self.shard_align = max(8, int(shard_align))
        self.shadow_skip = (int(shadow_skip[0]), int(shadow_skip[1]))
        self.rank, self.world = rank, world
        self.device = torch.device(device)
        self.server_mode = server_mode
        self.opt = server_opt or ServerOptConfig()
        self.upload_mode = upload
        self.pg = process_group
        if data_plane == "auto":
            data_plane = "native" if self.device.type == "cuda" else "collective"
        self.data_plane = data_plane
        self.n = (n_params + 7) // 8 * 8
        self.n_real = n_params
        self.epoch = 0
        self.server_step = 0
        self.last_status = 0
        # bit p set: rank p is alive. Reducers wait for every live rank to ARRIVE at the round
        # (a weight-0 node is alive but not reporting: it is still waited for -- its buffer must not
        # be overwritten while it is in use -- and still receives the new global model). A dead
        # node (mark_dead) is neither waited for nor written to.
        self.live_mask = (1 << world) - 1

        if data_plane == "native":
            assert self.device.type == "cuda"
            from ..ops import native

            self._C = native()
            self.heap = SymmetricHeap(rank, world, self.device)
            # "auto": the in-switch reduction / multicast store from 3 GPUs up.  Between two GPUs plain P2P loads and stores
            # move the same bytes faster (measured, profiles/comm_2gpu_r1c.json: K2 0.186 ms P2P vs 0.313 ms multicast for
            # 102 MB); at 8 it is the other way round (0.260 vs 0.316 ms, comm_8gpu_r1a.json).
            want_mc = (multicast is True) or (multicast == "auto" and world >= 3)
            self._w_buf = self.heap.alloc(self.n * 4, multicast=want_mc)
            self.w = self._w_buf.view(torch.float32, self.n)
            if upload == "weights_f32":
                self._up_buf = self._w_buf
                self.upload = self.w
            else:
                esz = 4 if upload == "delta_f32" else 2
                self._up_buf = self.heap.alloc(self.n * esz, multicast=want_mc)
                self.upload = self._up_buf.view(torch.float32 if esz == 4 else torch.bfloat16, self.n)
            self._shadow_buf = self.heap.alloc(self.n * 2, multicast=want_mc or shadow_multicast) if shadow_bf16 else None
            self.shadow = self._shadow_buf.view(torch.bfloat16, self.n) if shadow_bf16 else None
            self._pad_buf = self.heap.alloc(PAD_WORDS * 4, multicast=False)
            self.pad = self._pad_buf.view(torch.int32, PAD_WORDS)
            self.use_multicast = bool(self._w_buf.mc_ptr) and bool(self._up_buf.mc_ptr) and want_mc
            self._cta_counter = torch.zeros(4, dtype=torch.int32, device=self.device)
            self.timeout_cycles = int(timeout_ms * device_clock_khz(self.device))
            # K1: the current round as a device word (a captured first-step graph reads it) + a status word its waits report to
            self.k1_words = torch.zeros(4, dtype=torch.int32, device=self.device)
            self._k1_flag_bufs: list = []
        else:
            self.heap = None
            self.w = torch.zeros(self.n, dtype=torch.float32, device=self.device)
            if upload == "weights_f32":
                self.upload = self.w
            else:
                self.upload = torch.zeros(self.n, dtype=torch.float32 if upload == "delta_f32" else torch.bfloat16,
                                          device=self.device)
            self.shadow = torch.zeros(self.n, dtype=torch.bfloat16, device=self.device) if shadow_bf16 else None
            self.use_multicast = False
        # server state (fp32 master + optimizer moments). Full-size for simplicity; only the
        # owned slice is touched.  `reducers` = the ranks that own a slice of the global model: rank 0 alone in
        # `central` mode (the vantage6 central server), every live rank in `sharded` mode.
        self.reducers: List[int] = [0] if server_mode == "central" else list(range(world))
        self.w_global = torch.zeros(8, dtype=torch.float32, device=self.device)
        self.opt_m = torch.zeros(8, dtype=torch.float32, device=self.device)
        self.opt_v = torch.zeros(8, dtype=torch.float32, device=self.device)
        self._reshard()
        self.sm_count = (torch.cuda.get_device_properties(self.device).multi_processor_count
                         if self.device.type == "cuda" else 1)

    def _reshard(self) -> None:
        """(Re)compute this rank's slice [lo, hi) from the reducer list and make sure the server state it needs exists."""
        nr = len(self.reducers)
        pos = self.reducers.index(self.rank) if self.rank in self.reducers else None
        self.lo, self.hi, self._chunk = slice_bounds(self.n, nr, pos, self.shard_align)
        self.n_reducers = nr
        self.reducer_mask = sum(1 << r for r in self.reducers)
        if self.is_reducer:
            def grow(t: torch.Tensor, need: bool) -> torch.Tensor:
                if need and t.numel() < self.n:
                    return torch.zeros(self.n, dtype=torch.float32, device=self.device)
                return t
            self.w_global = grow(self.w_global, True)
            self.opt_m = grow(self.opt_m, self.opt.name in ("fedavgm", "fedadam"))
            self.opt_v = grow(self.opt_v, self.opt.name == "fedadam")

    # ------------------------------------------------------------------ K1: broadcast fused with the first consuming GEMM
    def k1_layer(self, offset: int, n_out: int, k_in: int) -> Optional[dict]:
        """Everything ``ops.gemm.bcast_push_gemm_bf16`` needs to deliver the [n_out, k_in] bf16 weight at flat ``offset`` (inside
        ``shadow_skip``): the multicast address of the shadow slice, symmetric per-tile ready flags, and the 256-row blocks
        of the layer that lie in THIS rank's reducer slice (their owner multicasts them).  ``None`` when the layout does not
        allow it (no multicast binding, slice boundaries that cut a 256-row block)."""
        if self.data_plane != "native" or self._shadow_buf is None or not self._shadow_buf.mc_ptr or self.world < 2:
            return None
        if n_out % 256 or offset % 8 or not (self.shadow_skip[0] <= offset and offset + n_out * k_in <= self.shadow_skip[1]):
            return None
        # the same answer on every rank (the flag allocation below is collective): NO reducer boundary may cut a block
        own = k1_own_blocks(self.lo, self.hi, self._chunk, len(self.reducers), offset, n_out, k_in)
        if own is None:
            return None
        n_flags = (n_out // 256) * ((k_in + 63) // 64)
        fb = self.heap.alloc(n_flags * 4, multicast=False)
        fb.view(torch.int32, n_flags).zero_()
        self._k1_flag_bufs.append(fb)
        return dict(w_mc_ptr=self._shadow_buf.mc_ptr + 2 * offset, flags=fb.view(torch.int32, n_flags), flag_peer_ptrs=list(fb.peer_ptrs),
                    own_blocks=own, is_owner=own[1] > own[0], world=self.world, epoch_ptr=self.k1_words.data_ptr(),
                    status_ptr=self.k1_words.data_ptr() + 4)

    def k1_status(self) -> int:
        """1 when a K1 tile never arrived (dead owner) since the last call; cleared."""
        if self.data_plane != "native":
            return 0
        st = int(self.k1_words[1].item())
        if st:
            self.k1_words[1:2].zero_()
        return st

    # ------------------------------------------------------------------ helpers
    @property
    def is_reducer(self) -> bool:
        return self.rank in self.reducers

    def nvlink_bytes_per_round(self) -> int:
        """Bytes this rank must pull + push over NVLink in one native round (roofline numerator)."""
        esz = 2 if self.upload_mode == "delta_bf16" else 4
        mine = self.hi - self.lo
        if self.world == 1:
            return 0
        pull = mine * esz * (1 if self.use_multicast else self.world - 1)
        push = mine * 4 * (1 if self.use_multicast else self.world - 1)
        return pull + push

    def _weights_vector(self, weight: Sequence[float]) -> List[float]:
        """An explicit per-rank vector (every rank must pass the same one)."""
        w = [float(x) for x in weight]
        assert len(w) == self.world
        return w

    # ------------------------------------------------------------------ rounds
    @torch.no_grad()
    def initialize_global(self) -> None:
        """Round 0: the model in rank 0's ``w`` becomes the global model on every node (one aggregation in which only
        rank 0 reports its *weights*, plain FedAvg with server lr 1, P2P path -- selected by the ``round0`` mode flag of
        :meth:`_aggregate`, the engine's configuration is not touched)."""
        self._aggregate([1.0] + [0.0] * (self.world - 1), count_step=False, force_p2p=True, round0=True)
        if self.shadow is not None:
            self.shadow.copy_(self.w.to(torch.bfloat16))

    @torch.no_grad()
    def aggregate(self, weight: float | Sequence[float] = 1.0, prescaled: bool = False) -> None:
        """One federated aggregation: after it returns (stream order) ``w`` is the new global.

        ``weight`` is either THIS rank's sample count ``n_i`` (a scalar: the usual case -- a node only knows its own
        data; the counts travel to the reducers next to the contributions, through the signal pads on the native
        plane and a tiny all-gather on the collective plane) or the full per-rank vector (identical on every rank;
        0 = that rank does not report this round).  ``prescaled``: the contribution buffer already holds
        ``n_i * delta`` (delta upload modes with the fused publish) -- unequal ``n_i`` then still take the in-switch
        ``multimem.ld_reduce`` path."""
        if isinstance(weight, (int, float)):
            self._aggregate(None, count_step=True, my_weight=float(weight), prescaled=prescaled)
        else:
            self._aggregate(self._weights_vector(weight), count_step=True, prescaled=prescaled)

    def mark_dead(self, rank: int, master_hint: torch.Tensor | None = None) -> None:
        """Exclude a failed node from all future rounds (every surviving rank must call this with
        the same argument, e.g. after ``poll_status()`` reported a timeout): the FedAvg weights
        renormalise over the reporters, as with any partial participation.  If the dead rank owned a slice of the
        global model the slices are re-partitioned over the surviving reducers (``central`` mode: the lowest live rank
        takes over as the server); the fp32 master of a newly owned range is seeded from ``master_hint`` (the global
        model of the round start, when the caller has it) or from the current parameters, and its server-optimizer
        moments restart from zero."""
        if not (self.live_mask >> rank) & 1:
            return
        self.live_mask &= ~(1 << rank)
        if rank in self.reducers:
            old_lo, old_hi = self.lo, self.hi
            live = [r for r in range(self.world) if (self.live_mask >> r) & 1]
            assert live, "no live rank left"
            self.reducers = [live[0]] if self.server_mode == "central" else [r for r in self.reducers if r != rank]
            self._reshard()
            if self.is_reducer and self.hi > self.lo:
                src = master_hint if master_hint is not None else self.w
                for a, b in ((self.lo, min(self.hi, old_lo)), (max(self.lo, old_hi), self.hi)):
                    if b > a:       # newly owned range
                        self.w_global[a:b].copy_(src[a:b])
                        if self.opt_m.numel() >= self.n:
                            self.opt_m[a:b].zero_()
                        if self.opt_v.numel() >= self.n:
                            self.opt_v[a:b].zero_()

    def _gather_weights(self, my_weight: float) -> List[float]:
        """Collective plane: exchange the per-rank sample counts (one tiny all-gather)."""
        if self.world == 1:
            return [my_weight]
        import torch.distributed as dist

        mine = torch.tensor([my_weight], dtype=torch.float64, device=self.device if self.device.type == "cuda" else "cpu")
        out = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(out, mine, group=self.pg)
        return [float(t.item()) for t in out]

    def _aggregate(self, weights: List[float] | None, count_step: bool, force_p2p: bool = False, my_weight: float = 0.0,
                   prescaled: bool = False, round0: bool = False) -> None:
        dynamic = weights is None
        opt = ServerOptConfig("fedavg", 1.0) if round0 else self.opt
        upload_mode = "weights_f32" if round0 else self.upload_mode
        if dynamic and self.data_plane != "native":
            weights, dynamic = self._gather_weights(my_weight), False
        if not dynamic:
            weights = [w if (self.live_mask >> r) & 1 else 0.0 for r, w in enumerate(weights)]
            total = sum(weights)
            assert total > 0, "at least one node must report"
        self.epoch += 1
        if count_step:
            self.server_step += 1
        t = max(self.server_step, 1)
        bias1 = 1.0 / (1.0 - opt.beta1 ** t)
        bias2 = 1.0 / (1.0 - opt.beta2 ** t)
        if self.data_plane == "native":
            mode_idx = UPLOAD_MODES.index(upload_mode)
            is_delta = mode_idx != 0
            # The kernel decides at run time (from the n_i it received) whether the in-switch reduction applies: every
            # rank reports and (contributions pre-scaled by n_i, or all n_i equal -> mean = sum / world); otherwise the
            # reducer applies n_i itself on the P2P path.  Multicast needs every bound GPU alive.
            all_live = self.live_mask == (1 << self.world) - 1
            mc_ok = self.use_multicast and not force_p2p and all_live
            from ..ops import stream_ptr

            wb = self._w_buf
            up = wb if round0 else self._up_buf           # round 0 reads the weights themselves
            null8 = [0] * 8
            self._C.fedavg_round(
                up.peer(), wb.peer(), self._shadow_buf.peer() if self._shadow_buf else null8, self._pad_buf.peer(),
                up.mc() if mc_ok else 0, wb.mc() if mc_ok else 0,
                self._shadow_buf.mc() if (self._shadow_buf and mc_ok) else 0,
                self.w_global.data_ptr(), self.opt_m.data_ptr(), self.opt_v.data_ptr(), weights if not dynamic else [0.0] * self.world,
                self.lo, self.hi, self.rank, self.world, self.n_reducers, self.live_mask, self.epoch, is_delta, bool(prescaled),
                SERVER_OPTS[opt.name], opt.lr, opt.beta1, opt.beta2, opt.eps, bias1, bias2,
                0.0, self.timeout_cycles, self._cta_counter.data_ptr(),
                1 if upload_mode == "delta_bf16" else 0,
                self.sm_count if (self.is_reducer and self.hi > self.lo) else 1, stream_ptr(),
                dynamic, float(my_weight), self.reducer_mask, self.shadow_skip[0], self.shadow_skip[1])
            if self.shadow_skip[1] > self.shadow_skip[0]:
                self.k1_words[0:1].fill_(self.epoch)          # stream-ordered after K2: the round K1's tile flags are compared with
        else:
            if prescaled:           # the collective arm applies the weights itself
                weights_eff = [1.0 if w > 0 else 0.0 for w in weights]
                self._aggregate_collective(weights_eff, total, bias1, bias2, opt, upload_mode)
            else:
                self._aggregate_collective(weights, total, bias1, bias2, opt, upload_mode)

    # -- baseline / CPU data plane -------------------------------------------------------------
    def _aggregate_collective(self, weights, total, bias1, bias2, opt=None, upload_mode=None) -> None:
        import torch.distributed as dist

        opt = opt or self.opt
        upload_mode = upload_mode or self.upload_mode
        is_delta = upload_mode != "weights_f32"
        my_w = weights[self.rank]
        if my_w == 0:
            contrib = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        elif is_delta:
            contrib = self.upload.float() * my_w
        else:
            contrib = self.w * my_w
        if self.world > 1:
            if self.server_mode == "central":
                dist.reduce(contrib, dst=self.reducers[0], group=self.pg)
            else:
                dist.all_reduce(contrib, group=self.pg)
        if self.is_reducer:
            lo, hi = (0, self.n) if self.server_mode == "central" else (self.lo, self.hi)
            mean = contrib[lo:hi] / total
            wg = self.w_global[lo:hi]
            d = mean if is_delta else mean - wg
            o = opt
            if o.name == "fedavgm":
                m = self.opt_m[lo:hi]
                m.mul_(o.beta1).add_(d)
                wg.add_(m, alpha=o.lr)
            elif o.name == "fedadam":
                m, v = self.opt_m[lo:hi], self.opt_v[lo:hi]
                m.mul_(o.beta1).add_(d, alpha=1 - o.beta1)
                v.mul_(o.beta2).addcmul_(d, d, value=1 - o.beta2)
                wg.add_((m * bias1) / ((v * bias2).sqrt() + o.eps), alpha=o.lr)
            else:
                wg.add_(d, alpha=o.lr)
        if self.server_mode == "central":
            if self.rank == self.reducers[0]:
                self.w.copy_(self.w_global)
            if self.world > 1:
                dist.broadcast(self.w, src=self.reducers[0], group=self.pg)
        else:
            if self.world > 1:
                out = torch.zeros_like(self.w)
                out[self.lo:self.hi] = self.w_global[self.lo:self.hi]
                dist.all_reduce(out, group=self.pg)
                self.w.copy_(out)
            else:
                self.w.copy_(self.w_global)
        if self.shadow is not None:
            self.shadow.copy_(self.w.to(torch.bfloat16))

    # ------------------------------------------------------------------ fault handling
    def poll_status(self) -> int:
        """Non-zero when a kernel wait timed out (a peer died / was stopped mid-round)."""
        if self.data_plane != "native":
            return 0
        from ..ops import native

        self.last_status = int(self.pad[native().PAD_STATUS].item())
        if self.last_status == 0 and self.shadow_skip[1] > self.shadow_skip[0] and self.k1_status():
            self.last_status = 3                     # a K1 weight tile never arrived (its owner died during the first local step)
        return self.last_status

    def missing_mask(self) -> int:
        """Bitmask of the ranks whose contribution (or slice push) never arrived in a failed round: this rank's own
        waits OR'd with what the other reducers reported (csrc/fedavg.cu step 6)."""
        if self.data_plane != "native":
            return 0
        from ..ops import native

        return int(self.pad[native().PAD_MISSING].item()) & ((1 << self.world) - 1)

    def clear_status(self) -> None:
        if self.data_plane == "native":
            from ..ops import native

            self.pad[native().PAD_STATUS] = 0
            self.pad[native().PAD_MISSING] = 0
            self.last_status = 0

    def abort(self) -> None:
        """Raise the abort flag on this rank's pad: every kernel waiting on it returns promptly."""
        if self.data_plane == "native":
            from ..ops import native

            self.pad[native().PAD_ABORT] = 1

    # ------------------------------------------------------------------ checkpoint
    def state_dict(self) -> dict:
        return {"w_global": self.w_global, "opt_m": self.opt_m, "opt_v": self.opt_v, "server_step": self.server_step,
                "lo": self.lo, "hi": self.hi, "n": self.n}

    def load_state_dict(self, sd: dict) -> None:
        self.w_global.copy_(sd["w_global"])
        self.opt_m.copy_(sd["opt_m"])
        self.opt_v.copy_(sd["opt_v"])
        self.server_step = int(sd["server_step"])

    def close(self) -> None:
        if self.heap is not None:
            self.w = self.upload = self.shadow = self.pad = None  # type: ignore[assignment]
            self.heap.close()
            self.heap = None


# ----------------------------------------------------------------------------------------------
# K3: small-message aggregation (GLM coefficients, the 1k-parameter vector)
# ----------------------------------------------------------------------------------------------
class SmallAggregator:
    """One-shot weighted all-reduce for payloads <= 64 KB (latency-bound path, csrc K3)."""

    MAX_FLOATS = 16384

    def __init__(self, n_floats: int, rank: int = 0, world: int = 1, device="cpu", process_group=None,
                 timeout_ms: float = 20000.0):
        assert n_floats <= self.MAX_FLOATS
        self.n = (n_floats + 3) // 4 * 4
        self.rank, self.world, self.device = rank, world, torch.device(device)
        self.pg = process_group
        self.epoch = 0
        self.native = self.device.type == "cuda"
        if self.native:
            from ..ops import native

            self._C = native()
            self.heap = SymmetricHeap(rank, world, self.device)
            self._slots = self.heap.alloc(2 * self.n * 4, multicast=False)      # double-buffered by epoch parity
            self._pad_buf = self.heap.alloc(PAD_WORDS * 4, multicast=False)
            self.slots = self._slots.view(torch.float32, 2 * self.n).view(2, self.n)
            self.timeout_cycles = int(timeout_ms * device_clock_khz(self.device))
        else:
            self.heap = None
            self.slots = torch.zeros(2, self.n, dtype=torch.float32)
        self.out = torch.zeros(self.n, dtype=torch.float32, device=self.device)

    def slot(self) -> torch.Tensor:
        """The buffer the caller fills with its payload for the NEXT all-reduce."""
        return self.slots[(self.epoch + 1) & 1]

    @torch.no_grad()
    def allreduce(self, weights: float | Sequence[float] = 1.0, normalize: bool = True) -> torch.Tensor:
        w = [float(weights)] * self.world if isinstance(weights, (int, float)) else [float(x) for x in weights]
        total = sum(w) if normalize else 1.0
        self.epoch += 1
        par = self.epoch & 1
        if self.native:
            from ..ops import count, stream_ptr

            off = par * self.n * 4
            count(1)
            self._C.small_allreduce(self._slots.peer(off), self._pad_buf.peer(), w, self.out.data_ptr(), self.n,
                                    self.rank, self.world, self.epoch, 1.0 / total, self.timeout_cycles, stream_ptr())
        else:
            import torch.distributed as dist

            contrib = self.slots[par] * w[self.rank]
            if self.world > 1:
                dist.all_reduce(contrib, group=self.pg)
            self.out.copy_(contrib / total)
        return self.out

    def close(self):
        if self.heap is not None:
            self.slots = None  # type: ignore[assignment]
            self.heap.close()
            self.heap = None
