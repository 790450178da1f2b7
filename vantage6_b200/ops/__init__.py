"""Native sm_100a operators of vantage6_b200.

``native()`` returns the compiled extension (``_C``), building it in-tree when it is missing
and ``nvcc`` is available.  On a machine WITH a GPU a missing extension is a hard error (no
silent eager fallback -- the CUDA path must be the one that runs); on a CPU-only machine the
Python reference implementations in each module are used by the ``not gpu`` tests.

Kernel inventory:  K1 ``gemm.bcast_linear``, K2 ``fedavg.fedavg_round``,
K3 ``fedavg.small_allreduce``, K5 ``norm``, K6 ``rope``, K7 ``optim``, K8 ``glm``,
tcgen05 GEMM ``gemm.linear``.
"""
from __future__ import annotations

import importlib
import os
import threading

_lock = threading.Lock()
_C = None
_load_error: Exception | None = None


def native(required: bool = True):
    """Return the native module, importing (and if necessary building) it once."""
    global _C, _load_error
    if _C is not None:
        return _C
    with _lock:
        if _C is not None:
            return _C
        try:
            _C = importlib.import_module("vantage6_b200.ops._C")
        except ImportError as first:
            try:
                from . import build as _build

                if os.environ.get("V6B200_NO_BUILD") == "1":
                    raise first
                _build.build(verbose=False)
                importlib.invalidate_caches()
                _C = importlib.import_module("vantage6_b200.ops._C")
            except Exception as e:  # noqa: BLE001
                _load_error = e
                if required:
                    raise RuntimeError(
                        "vantage6_b200 native extension (ops/_C*.so) is not built and could not be built: "
                        f"{e!r}. Run `python -m vantage6_b200.ops.build`.") from e
                return None
    return _C


def have_native() -> bool:
    return native(required=False) is not None


def cuda_ready() -> bool:
    """True when torch sees a GPU *and* the native extension is loadable."""
    import torch

    if not torch.cuda.is_available():
        return False
    native(required=True)   # loud failure on a GPU box
    return True


# Number of launches of OUR kernels issued from Python (bench.py `gpu_launches`): wrappers call
# count(n) next to each native call; kernels replayed inside a CUDA graph are accounted for by the
# trainer (launches recorded at capture time x number of replays).
LAUNCHES = [0]


def count(n: int = 1) -> None:
    LAUNCHES[0] += n


def stream_ptr() -> int:
    import torch

    return torch.cuda.current_stream().cuda_stream


_tree_scratch: dict = {}


def tree_scratch(device):
    """Scratch of the deterministic in-kernel tree reductions (csrc/tree_reduce.cuh): per-CTA partials plus
    self-resetting arrival counters, zeroed once, shared by every BatchNorm / bias-backward launch of the device
    (launches are stream-ordered)."""
    import torch

    key = str(device)
    buf = _tree_scratch.get(key)
    if buf is None:
        buf = _tree_scratch[key] = torch.zeros(int(native().BN_SCRATCH_FLOATS), device=device, dtype=torch.float32)
    return buf
