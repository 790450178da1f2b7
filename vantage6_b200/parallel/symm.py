"""Symmetric heap: the B200 data plane of the federation.

One process per GPU.  ``SymmetricHeap.alloc`` is a collective that gives every rank

* ``buf.local``      -- a torch tensor view of its own allocation,
* ``buf.peer_ptrs``  -- device addresses of *every* rank's allocation mapped into this process
                        (plain loads/stores on them travel over NVLink 5 / NVSwitch),
* ``buf.mc_ptr``     -- an NVLS multicast address over all of them (0 when unsupported):
                        ``multimem.st`` broadcasts in the switch, ``multimem.ld_reduce`` reduces in it.

Backends
--------
``native`` (default): csrc/symm.cpp -- CUDA VMM (cuMemCreate + POSIX fd) with fd passing over a
Unix-socket mesh; no NCCL / NVSHMEM involved.
``local``: world_size == 1 or CPU -- ordinary tensors (used by the CPU tests; kernels are
replaced by their PyTorch references).

This replaces the reference's data plane, where task inputs/results travel as (optionally RSA
encrypted) blobs inside REST payloads through a SQL database (reference
vantage6/cli/node.py:591-614, vantage6/cli/context.py:30-42).
"""
from __future__ import annotations

import os
import tempfile
from dataclasses import dataclass, field
from typing import List, Optional

import torch

PAD_WORDS = 256          # uint32 words per rank signal pad (1 KB)


class _CudaPtr:
    """Adapter exposing a raw device pointer through ``__cuda_array_interface__``."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {
            "shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None,
        }


def tensor_from_ptr(ptr: int, nbytes: int, device: torch.device) -> torch.Tensor:
    return torch.as_tensor(_CudaPtr(ptr, nbytes), device=device)


@dataclass
class SymmBuffer:
    local: torch.Tensor                 # uint8 view of this rank's allocation
    peer_ptrs: List[int]
    mc_ptr: int
    nbytes: int
    aid: int = -1
    keepalive: list = field(default_factory=list)

    def view(self, dtype: torch.dtype, numel: Optional[int] = None) -> torch.Tensor:
        t = self.local.view(dtype)
        return t if numel is None else t[:numel]

    def peer(self, offset_bytes: int = 0) -> List[int]:
        return [p + offset_bytes if p else 0 for p in self.peer_ptrs]

    def mc(self, offset_bytes: int = 0) -> int:
        return self.mc_ptr + offset_bytes if self.mc_ptr else 0


class SymmetricHeap:
    _seq = 0

    def __init__(self, rank: int, world: int, device: torch.device | int | str = "cpu",
                 rendezvous_dir: Optional[str] = None, timeout_s: int = 120):
        self.rank, self.world = rank, world
        self.device = torch.device(device) if not isinstance(device, torch.device) else device
        self.native = self.device.type == "cuda"
        self.gid = -1
        self.multicast = False
        if self.native:
            from ..ops import native

            self._C = native()
            if rendezvous_dir is None:
                tag = os.environ.get("MASTER_PORT", "0")
                rendezvous_dir = os.path.join(tempfile.gettempdir(), f"v6b200_symm_{os.getuid()}_{tag}_{SymmetricHeap._seq}")
            SymmetricHeap._seq += 1
            os.makedirs(rendezvous_dir, exist_ok=True)
            self.dir = rendezvous_dir
            dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
            if torch.cuda.current_device() != dev_index:
                # our kernels launch on the CURRENT device's stream: a process that owns this heap must run on its GPU
                # (nodes are pinned with V6_GPU while every GPU stays visible so that peers can be mapped)
                torch.cuda.set_device(dev_index)
            self.gid = self._C.symm_init(rank, world, dev_index, rendezvous_dir, timeout_s)
            self.multicast = bool(self._C.symm_multicast_supported(self.gid)) and world > 1
        else:
            assert world == 1, "the CPU ('local') symmetric heap only supports world_size == 1"
        self.buffers: List[SymmBuffer] = []

    def alloc(self, nbytes: int, multicast: bool = True) -> SymmBuffer:
        nbytes = (nbytes + 15) // 16 * 16
        if self.native:
            aid, ptrs, mc, padded = self._C.symm_alloc(self.gid, nbytes, bool(multicast and self.multicast))
            local = tensor_from_ptr(ptrs[self.rank], padded, self.device)
            buf = SymmBuffer(local=local, peer_ptrs=list(ptrs[: self.world]), mc_ptr=int(mc), nbytes=padded, aid=aid)
        else:
            local = torch.zeros(nbytes, dtype=torch.uint8)
            buf = SymmBuffer(local=local, peer_ptrs=[local.data_ptr()], mc_ptr=0, nbytes=nbytes)
        self.buffers.append(buf)
        return buf

    def host_barrier(self) -> None:
        if self.native and self.world > 1:
            self._C.symm_barrier_host(self.gid)

    def close(self) -> None:
        if self.native and self.gid >= 0:
            torch.cuda.synchronize(self.device)
            for b in self.buffers:
                b.local = None  # type: ignore[assignment]
            self._C.symm_finalize(self.gid)
            self.gid = -1

    def __del__(self):  # best effort
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass
