"""Node-to-node channel for algorithms -- the in-box equivalent of vantage6's algorithm VPN.

The reference provisions an optional VPN so that algorithm containers of different nodes can talk to each other
directly instead of through the central server (reference vantage6/cli/configuration_wizard.py:78-84 ``vpn_subnet``,
:159-195 ``vpn_server``, vantage6/cli/context.py:133-138 the per-node VPN volume).  On an NVSwitch box the direct path
between nodes is not a tunnel but memory: this module gives an algorithm that API.

    from vantage6_b200.algorithm.peer import PeerChannel

    ch = PeerChannel.open(rendezvous)            # rendezvous = {"addr", "port", "world", "ranks": {org_id: rank}} from the master
    buf = ch.alloc(n_bytes)                      # symmetric buffer: buf.local (this node's tensor), buf.peer_ptrs, buf.mc_ptr
    ch.barrier()
    total = ch.allreduce(local_vector, weight)   # weighted sum / mean over the nodes (<= 64 KB: the K3 latency path)
    ch.close()

GPU nodes (``V6_GPU`` set, CUDA available): NVLink symmetric heap (parallel/symm.py: CUDA VMM + fd passing, NVLS multicast
when the switch offers it) and the single-CTA small all-reduce kernel (csrc/fedavg.cu K3) -- no NCCL, no TCP.
CPU nodes: a gloo process group with the same API (what the CPU plumbing tests run).

What may cross this channel is the algorithm's business (model parameters, sufficient statistics); the node's database
stays in the node's process -- same rule as for the VPN in vantage6.
"""
from __future__ import annotations

import os
from typing import Optional, Sequence


class PeerChannel:
    def __init__(self, rank: int, world: int, device, heap=None, pg=None):
        self.rank, self.world, self.device = rank, world, device
        self.heap, self.pg = heap, pg
        self._aggs: dict = {}
        self._barrier_epoch = 0
        self._pad = None

    # ------------------------------------------------------------------ construction
    @classmethod
    def open(cls, rendezvous: Optional[dict] = None, organization_id: Optional[int] = None) -> "PeerChannel":
        import torch

        org = int(os.environ.get("V6_ORGANIZATION_ID", "0")) if organization_id is None else int(organization_id)
        rv = rendezvous or {"addr": "127.0.0.1", "port": 0, "world": 1, "ranks": {str(org): 0}}
        rank, world = int(rv["ranks"][str(org)]), int(rv["world"])
        if torch.cuda.is_available():
            from ..parallel.symm import SymmetricHeap

            dev = torch.device("cuda", int(os.environ.get("V6_GPU", "0")))
            torch.cuda.set_device(dev)
            os.environ["MASTER_PORT"] = str(rv.get("port", 0))          # rendezvous directory key of the heap
            return cls(rank, world, dev, heap=SymmetricHeap(rank, world, dev))
        pg = None
        if world > 1:
            import torch.distributed as dist

            if not dist.is_initialized():
                dist.init_process_group("gloo", init_method=f"tcp://{rv['addr']}:{rv['port']}", rank=rank, world_size=world)
            pg = dist.group.WORLD
        return cls(rank, world, torch.device("cpu"), pg=pg)

    # ------------------------------------------------------------------ memory
    def alloc(self, nbytes: int, multicast: bool = True):
        """A buffer every node can address: ``.local`` (uint8 tensor of this node), ``.peer_ptrs`` (device addresses of every
        node's copy, plain loads / stores on them travel over NVLink), ``.mc_ptr`` (NVLS multicast address or 0)."""
        if self.heap is not None:
            return self.heap.alloc(nbytes, multicast=multicast)
        import torch

        from ..parallel.symm import SymmBuffer

        local = torch.zeros((nbytes + 15) // 16 * 16, dtype=torch.uint8)
        return SymmBuffer(local=local, peer_ptrs=[0] * self.world, mc_ptr=0, nbytes=local.numel())

    # ------------------------------------------------------------------ collectives
    def barrier(self) -> None:
        if self.heap is not None:
            import torch

            torch.cuda.synchronize(self.device)
            self.heap.host_barrier()
        elif self.pg is not None:
            import torch.distributed as dist

            dist.barrier(group=self.pg)

    def allreduce(self, vec, weight: float | Sequence[float] = 1.0, normalize: bool = True):
        """Weighted sum (``normalize=False``) or weighted mean of a small float vector over the nodes; returns a tensor on
        this node's device.  ``weight`` is the per-node vector (identical on all nodes) or one scalar for everybody."""
        import torch

        from ..parallel.fedavg import SmallAggregator

        v = torch.as_tensor(vec, dtype=torch.float32).flatten()
        n = int(v.numel())
        agg = self._aggs.get(n)
        if agg is None:
            agg = self._aggs[n] = SmallAggregator(n, self.rank, self.world, self.device, process_group=self.pg)
        agg.slot()[:n].copy_(v.to(agg.slot().device))
        return agg.allreduce(weight, normalize=normalize)[:n].clone()

    def close(self) -> None:
        for a in self._aggs.values():
            a.close()
        self._aggs.clear()
        if self.heap is not None:
            self.heap.close()
            self.heap = None
