"""Long-lived GPU worker of a node (``vnode start --gpu K``).

The reference's node is a long-running service (reference vantage6/cli/node.py:380-410: ``vnode-local start`` inside a
container that stays up and launches one algorithm container per task).  The algorithm processes stay short-lived here
too -- but everything that is expensive on a GPU does NOT live in them: this worker owns the CUDA context, the NVLink
symmetric heap, the flat-buffer model, the fused optimizer state and the captured CUDA graphs of every (model,
federation) it has trained, and keeps them across tasks.  A ``fedavg`` task then only binds the node's data and runs
rounds: round 1's 12 s per-task bring-up (context + heap + model + graph capture, profiles/demo_network_2gpu_r1c.jsonl)
is paid once per node lifetime.

Protocol: newline-delimited JSON over a Unix socket (path in ``V6_GPU_WORKER``), one request at a time:

    {"op": "train", "kwargs": {...RPC_train kwargs...}, "organization_id": 3, "database_uri": "..."}  -> {"ok": true, "result": {...}}
    {"op": "glm", ...}   {"op": "ping"}   {"op": "stats"}   {"op": "shutdown"}

Run as:  python -m vantage6_b200.node.gpu_worker --socket PATH --gpu K
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import socket
import sys
import threading
import time
import traceback
from typing import Any, Dict, Optional

log = logging.getLogger("gpu-worker")


class GpuWorker:
    def __init__(self, sock_path: str, gpu: Optional[int]):
        self.sock_path, self.gpu = sock_path, gpu
        self.trainers: Dict[tuple, Any] = {}        # federation key -> (trainer, spec): kept warm across tasks
        self.databases: Dict[str, Any] = {}
        self.stats = {"tasks": 0, "trainer_builds": 0, "trainer_reuses": 0, "started": time.time()}
        self._stop = threading.Event()

    # ------------------------------------------------------------------ warm-up
    def warm(self) -> None:
        import torch

        if torch.cuda.is_available():
            dev = torch.device("cuda", self.gpu or 0)
            torch.cuda.set_device(dev)
            torch.zeros(1, device=dev)              # CUDA context
            from ..ops import native

            native()                                # load the sm_100a extension once
            log.info("GPU worker warm on %s (%s)", dev, torch.cuda.get_device_name(dev))
        else:
            log.info("GPU worker running without a GPU (CPU data plane)")

    def database(self, uri: Optional[str]):
        from ..algorithm.wrapper import load_data

        if not uri:
            return None
        key = uri
        try:
            if "://" not in uri or uri.startswith("file://"):
                p = uri[7:] if uri.startswith("file://") else uri
                key = f"{uri}@{os.path.getmtime(p)}"
        except OSError:
            pass
        if key not in self.databases:
            self.databases.clear()                  # one resident database per node is enough
            self.databases[key] = load_data(uri)
        return self.databases[key]

    # ------------------------------------------------------------------ requests
    def handle(self, req: dict) -> dict:
        op = req.get("op")
        if op == "ping":
            return {"ok": True, "result": {"pid": os.getpid(), "gpu": self.gpu}}
        if op == "stats":
            return {"ok": True, "result": dict(self.stats, trainers=[list(map(str, k)) for k in self.trainers])}
        if op == "shutdown":
            self._stop.set()
            return {"ok": True, "result": None}
        if op == "train":
            from ..algorithm.builtin import fedavg

            os.environ["V6_ORGANIZATION_ID"] = str(req.get("organization_id", 0))
            data = self.database(req.get("database_uri"))
            self.stats["tasks"] += 1
            return {"ok": True, "result": fedavg.train_partial(data, trainer_cache=self, **(req.get("kwargs") or {}))}
        return {"ok": False, "error": f"unknown op {op!r}"}

    # trainer cache protocol used by fedavg.train_partial
    def get_trainer(self, key: tuple):
        hit = self.trainers.get(key)
        if hit is not None:
            self.stats["trainer_reuses"] += 1
            log.info("reusing the resident trainer %s: symmetric heap, model, optimizer and CUDA graphs stay as they are", key)
        return hit

    def put_trainer(self, key: tuple, value) -> None:
        self.trainers[key] = value
        self.stats["trainer_builds"] += 1
        log.info("built trainer %s (kept for the following tasks)", key)

    # ------------------------------------------------------------------ server loop
    def serve(self) -> None:
        try:
            os.unlink(self.sock_path)
        except FileNotFoundError:
            pass
        srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        srv.bind(self.sock_path)
        os.chmod(self.sock_path, 0o600)
        srv.listen(8)
        srv.settimeout(0.5)
        print(f"gpu-worker ready on {self.sock_path} (gpu={self.gpu})", flush=True)
        while not self._stop.is_set():
            try:
                conn, _ = srv.accept()
            except socket.timeout:
                continue
            with conn:
                conn.settimeout(None)
                f = conn.makefile("rwb")
                line = f.readline()
                if not line:
                    continue
                try:
                    rep = self.handle(json.loads(line))
                except Exception as e:  # noqa: BLE001
                    traceback.print_exc()
                    rep = {"ok": False, "error": repr(e), "traceback": traceback.format_exc()[-4000:]}
                f.write((json.dumps(rep) + "\n").encode())
                f.flush()
        srv.close()
        for tr, _ in self.trainers.values():
            try:
                tr.close()
            except Exception:  # noqa: BLE001
                pass
        try:
            os.unlink(self.sock_path)
        except OSError:
            pass


def call(sock_path: str, req: dict, timeout: Optional[float] = None) -> dict:
    """Client side (no torch import): send one request, wait for the reply."""
    s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    s.settimeout(timeout)
    s.connect(sock_path)
    with s:
        f = s.makefile("rwb")
        f.write((json.dumps(req) + "\n").encode())
        f.flush()
        line = f.readline()
    if not line:
        raise RuntimeError("GPU worker closed the connection without a reply")
    rep = json.loads(line)
    if not rep.get("ok"):
        raise RuntimeError(f"GPU worker error: {rep.get('error')}\n{rep.get('traceback', '')}")
    return rep["result"]


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--socket", required=True)
    ap.add_argument("--gpu", type=int, default=None)
    a = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO, stream=sys.stdout, format="%(asctime)s - %(name)-10s - %(levelname)-7s - %(message)s")
    if a.gpu is not None:
        os.environ["V6_GPU"] = str(a.gpu)
    w = GpuWorker(a.socket, a.gpu)
    w.warm()
    w.serve()
    return 0


if __name__ == "__main__":
    sys.exit(main())
