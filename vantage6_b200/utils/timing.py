"""Device-side timing, clock sampling and distributed helpers (the reference has
no profiling hooks at all; the north star demands device-timed, max-over-ranks numbers).
"""
from __future__ import annotations

import os
import statistics
import subprocess
import tempfile
from contextlib import contextmanager
from typing import Dict, Optional

import torch


class DeviceTimer:
    """CUDA-event stopwatch on the current stream (synchronises on both sides)."""

    def __init__(self, device=None):
        self.device = device
        self.start_ev = torch.cuda.Event(enable_timing=True)
        self.stop_ev = torch.cuda.Event(enable_timing=True)

    def start(self):
        torch.cuda.synchronize(self.device)
        self.start_ev.record()

    def stop(self) -> float:
        self.stop_ev.record()
        torch.cuda.synchronize(self.device)
        return self.start_ev.elapsed_time(self.stop_ev)   # milliseconds


def max_over_ranks(value: float, device=None) -> float:
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier_sync(device=None) -> None:
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    if device is not None and torch.device(device).type == "cuda":
        torch.cuda.synchronize(device)


_QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
          "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
          "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")


class ClockSampler:
    """Samples ``nvidia-smi`` clocks / throttle reasons every 200 ms DURING a timed region
    (B200_PROFILING.md "clocks line")."""

    def __init__(self, gpu_index: int = 0, period_ms: int = 200):
        self.gpu_index = gpu_index
        self.period_ms = period_ms
        self.proc: Optional[subprocess.Popen] = None
        self.path = os.path.join(tempfile.gettempdir(), f"v6b200_clocks_{os.getpid()}.csv")

    def __enter__(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={_QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu_index), "-lms", str(self.period_ms)],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001 -- no nvidia-smi (CPU box)
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:  # noqa: BLE001
                self.proc.kill()
            self.f.close()

    def summary(self) -> Dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        try:
            rows = [r.strip().split(",") for r in open(self.path) if r.strip()]
        except Exception:  # noqa: BLE001
            return out
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, val in zip(names, r[5:9]):
                    if val.strip().lower().startswith("active"):
                        reasons.add(nm)
            except Exception:  # noqa: BLE001
                continue
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


@contextmanager
def nvtx_range(name: str):
    """NVTX range when CUDA is available (visible in ncu / nsys timelines)."""
    if torch.cuda.is_available():
        torch.cuda.nvtx.range_push(name)
        try:
            yield
        finally:
            torch.cuda.nvtx.range_pop()
    else:
        yield
