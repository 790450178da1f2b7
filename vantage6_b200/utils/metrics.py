"""Per-round metrics as JSON lines under the instance log dir (the reference has
console breadcrumbs and a rotating file log but no metrics endpoint)."""
from __future__ import annotations

import json
import os
import time
from pathlib import Path
from typing import Any, Dict, Optional


class MetricsWriter:
    def __init__(self, path: Optional[os.PathLike] = None, static: Optional[Dict[str, Any]] = None):
        self.path = Path(path) if path else None
        self.static = dict(static or {})
        if self.path:
            self.path.parent.mkdir(parents=True, exist_ok=True)
        self.rows = []

    def log(self, **fields) -> Dict[str, Any]:
        row = {"ts": time.time(), **self.static, **fields}
        self.rows.append(row)
        if self.path:
            with open(self.path, "a") as f:
                f.write(json.dumps(row) + "\n")
        return row

    def round(self, idx: int, ms: float, loss: float, nvlink_bytes: int = 0, **extra) -> Dict[str, Any]:
        """One federated round: duration, mean local loss, achieved NVLink bus GB/s."""
        bus = (nvlink_bytes / (ms * 1e-3) / 1e9) if (ms and nvlink_bytes) else None
        return self.log(event="round", round=idx, ms=ms, rounds_per_sec=1e3 / ms if ms else None, loss=loss,
                        nvlink_bytes=nvlink_bytes, bus_GBps=bus, **extra)

    def log_round(self, idx: int, **fields) -> Dict[str, Any]:
        """Bookkeeping row written by ``FederatedTrainer.run_round`` (no device synchronisation: timings and the loss
        are added by whoever reads them back, e.g. the ``fedavg`` algorithm through :meth:`round`)."""
        return self.log(event=fields.pop("event", "round_done"), round=idx, **fields)
