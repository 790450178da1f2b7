"""Node runtime (``vnode-local start``: what the reference launches inside the node container,
reference vantage6/cli/node.py:380-382; behaviour per SURVEY.md Appendix C).

One node = one process pinned to one GPU.  It

1. authenticates at the central server with its ``api_key`` (-> node JWT),
2. sets up encryption (RSA private key vs the organization's public key at the server),
3. starts the local proxy server for algorithm -> server traffic,
4. syncs the results that were assigned while it was offline, then listens on the event
   channel for new tasks of its collaboration,
5. runs every assigned result's algorithm in an isolated child process with vantage6's
   input/output/token file + environment contract (algorithm/wrapper.py) inside a per-run
   temporary volume (``...-{run_id}-tmpvol``: reference vantage6/cli/context.py:140-141),
6. PATCHes started_at / finished_at / result / log back,
7. heartbeats (``last_seen``) and marks itself offline on shutdown.
"""
from __future__ import annotations

import json
import logging
import os
import queue
import re
import signal
import subprocess
import sys
import threading
import time
from pathlib import Path
from typing import Dict, Optional

from .. import __version__
from ..algorithm import resolve_image
from ..client import ClientBase, ServerError
from ..common import base64s_to_bytes, bytes_to_base64s
from ..common.encryption import DummyCryptor, RSACryptor
from ..runtime import runtime_dir
from .proxy import ProxyServer
from .zygote import Zygote, ZygoteProcess, preload_for

log = logging.getLogger("node")


def _now() -> str:
    import datetime as dt

    return dt.datetime.now(dt.timezone.utc).isoformat()


class NodeClient(ClientBase):
    """Server client with node (api_key) authentication."""

    def authenticate(self, api_key: str, gpu: Optional[int] = None) -> None:  # type: ignore[override]
        super().authenticate({"api_key": api_key, "gpu": gpu}, path="token/node")
        node = self.request(self._auth_reply["node_url"][len(self.path):])
        self.node_id = node["id"]
        self.name = node["name"]
        self.collaboration_id = node["collaboration"]["id"]
        self.organization_id = node["organization"]["id"]
        org = self.request(f"organization/{self.organization_id}")
        self.organization_name = org["name"]


class Node:
    def __init__(self, ctx, heartbeat_s: float = 15.0):
        self.ctx = ctx
        self.config = ctx.config
        self.heartbeat_s = heartbeat_s
        self.client = NodeClient(self.config["server_url"], self.config.get("port"), self.config.get("api_path", "/api"),
                                 ca_file=self.config.get("server_ca_file"))
        self.proxy = ProxyServer(self)
        self.queue: "queue.Queue[dict]" = queue.Queue()
        self.running: Dict[int, subprocess.Popen] = {}           # result id -> algorithm process
        self._task_of: Dict[int, int] = {}                          # result id -> task id (for kill requests)
        self.zygote: Optional[Zygote] = None                        # warm-start helper for algorithm runs
        self._seen: set = set()
        self._stop = threading.Event()
        self._threads = []
        self.cryptor = DummyCryptor()
        self.gpu = self._gpu_index()
        self._org_keys: Dict[int, Optional[str]] = {}
        self.gpu_worker: Optional[subprocess.Popen] = None          # resident GPU worker (node/gpu_worker.py)
        self.gpu_worker_sock: Optional[str] = None

    # ------------------------------------------------------------------ setup
    def _gpu_index(self) -> Optional[int]:
        pinned = os.environ.get("V6_GPU")
        if pinned not in (None, ""):
            try:
                return int(pinned)
            except ValueError:
                return None
        g = self.config.get("gpu") if hasattr(self.config, "get") else None
        return int(g) if g is not None else None

    def authenticate(self, retries: int = 60) -> None:
        for attempt in range(retries):
            try:
                self.client.authenticate(self.config["api_key"], self.gpu)
                log.info("Node '%s' authenticated (id=%s, organization=%s, collaboration=%s)", self.client.name,
                         self.client.node_id, self.client.organization_name, self.client.collaboration_id)
                return
            except ServerError as e:
                if e.status == 401:
                    raise
                log.warning("server not ready (%s); retrying", e)
            except Exception as e:  # noqa: BLE001 -- connection refused while the server starts
                log.warning("cannot reach server (%s); retry %d/%d", type(e).__name__, attempt + 1, retries)
            time.sleep(1.0)
        raise RuntimeError("could not authenticate with the server")

    def setup_encryption(self) -> None:
        collab = self.client.request(f"collaboration/{self.client.collaboration_id}")
        enc_cfg = self.config.get("encryption", {}) or {}
        if collab.get("encrypted") != bool(enc_cfg.get("enabled")):
            msg = (f"Expectations on encryption don't match! server: {collab.get('encrypted')}, "
                   f"node config: {bool(enc_cfg.get('enabled'))}")
            log.critical(msg)
            raise RuntimeError(msg)
        if not collab.get("encrypted"):
            log.warning("Disabling encryption!")
            self.cryptor = DummyCryptor()
            return
        key_file = os.environ.get("PRIVATE_KEY") or enc_cfg.get("private_key") or "private_key.pem"
        key_file = self.ctx.get_data_file(key_file)
        if not Path(key_file).exists():
            raise FileNotFoundError(f"private key {key_file} not found; run `vnode create-private-key`")
        self.cryptor = RSACryptor(key_file)
        org = self.client.request(f"organization/{self.client.organization_id}")
        if not org.get("public_key") or not self.cryptor.verify_public_key(org["public_key"]):
            log.warning("Local public key differs from the server's: uploading ours")
            self.client.request(f"organization/{self.client.organization_id}", method="patch",
                                json={"public_key": self.cryptor.public_key_str})

    # ------------------------------------------------------------------ crypto helpers (used by the proxy)
    def _org_public_key(self, org_id: int) -> Optional[str]:
        if org_id not in self._org_keys:
            self._org_keys[org_id] = self.client.request(f"organization/{org_id}").get("public_key")
        return self._org_keys[org_id]

    def encrypt_for_organization(self, plain: bytes, org_id: int) -> str:
        if isinstance(self.cryptor, RSACryptor):
            pub = self._org_public_key(org_id)
            if not pub:
                raise ValueError(f"organization {org_id} has no public key")
            return self.cryptor.encrypt_bytes_to_str(plain, pub)
        return self.cryptor.encrypt_bytes_to_str(plain, "")

    def decrypt_to_plain_b64(self, value: Optional[str]) -> Optional[str]:
        if not value:
            return value
        return bytes_to_base64s(self.cryptor.decrypt_str_to_bytes(value))

    # ------------------------------------------------------------------ task intake
    def sync_open_results(self) -> None:
        rows = self.client.request("result", params={"state": "open", "node_id": self.client.node_id, "include": "task"})
        for r in rows:
            self._enqueue(r)
        log.info("received %d open task(s) from the server", len(rows))

    def _enqueue(self, result: dict) -> None:
        if result["id"] in self._seen:
            return
        self._seen.add(result["id"])
        self.queue.put(result)

    def _handle_event(self, ev: dict) -> None:
        if ev["name"] == "new_task" and ev["data"].get("organization_id") == self.client.organization_id:
            if ev["data"].get("result_id") in self._seen:
                return
            # the event addressed to this node's own room carries the work item; the collaboration-wide one only names it
            r = ev["data"].get("result") or self.client.request(f"result/{ev['data']['result_id']}", params={"include": "task"})
            if r.get("finished_at") is None:
                self._enqueue(r)
        elif ev["name"] == "kill_containers":
            self.kill_task(ev["data"].get("task_id"))

    def _listen_websocket(self, since) -> Optional[int]:
        """Push channel (server/ws_events.py): returns the last event id seen when the connection ends, ``None`` when the
        server offers no websocket channel (-> long-poll)."""
        if os.environ.get("V6B200_EVENTS", "ws") != "ws":
            return None
        try:
            from websockets.sync.client import connect
        except Exception:  # noqa: BLE001
            return None
        port = self.client.request("health").get("event_port")
        if not port:
            return None
        from urllib.parse import urlsplit

        host = urlsplit(str(self.client.host)).hostname or "127.0.0.1"
        if since is None:
            since = self.client.request("health").get("events", 0)
        secure = urlsplit(str(self.client.host)).scheme == "https"
        url = f"{'wss' if secure else 'ws'}://{host}:{port}/?token={self.client.token}&since={since}"
        with connect(url, open_timeout=10, max_size=1 << 20, ssl=self.client._http.ssl_context() if secure else None) as ws:
            log.info("event channel: websocket %s:%s", host, port)
            self.sync_open_results()            # close the race between the first sync and the subscription
            while not self._stop.is_set():
                try:
                    msg = ws.recv(timeout=1.0)
                except TimeoutError:
                    continue
                ev = json.loads(msg)
                since = max(since, ev["id"])
                self._handle_event(ev)
        return since

    def _listen(self) -> None:
        since = None
        while not self._stop.is_set():
            try:
                got = self._listen_websocket(since)
                if got is not None:
                    since = got
                    continue
                params = {"timeout": 20}
                if since is not None:
                    params["since"] = since
                reply = self.client.request("event", params=params, timeout=40)
                for ev in reply.get("events", []):
                    self._handle_event(ev)
                if since is None:
                    self.sync_open_results()        # close the race between the first sync and the subscription
                since = reply.get("last_id", since)
            except Exception as e:  # noqa: BLE001
                if self._stop.is_set():
                    return
                log.warning("event channel error (%s); reconnecting", e)
                time.sleep(1.0)
                try:
                    self.sync_open_results()
                except Exception:  # noqa: BLE001
                    pass

    def _heartbeat(self) -> None:
        while not self._stop.wait(self.heartbeat_s):
            try:
                self.client.request(f"node/{self.client.node_id}", method="patch", json={"status": "online", "gpu": self.gpu})
            except Exception as e:  # noqa: BLE001
                log.debug("heartbeat failed: %s", e)

    # ------------------------------------------------------------------ execution
    def check_image_allowed(self, image: str) -> None:
        """Node-side policy (vantage6 node configuration ``allowed_images``): a list of regular expressions; when present,
        an algorithm image runs only if one of them matches its full name.  This is the data station's own veto -- it holds
        whatever the server or the researcher says."""
        allowed = self.config.get("allowed_images")
        if not allowed:
            return
        if isinstance(allowed, str):
            allowed = [allowed]
        if not any(re.fullmatch(str(rx), image) for rx in allowed):
            raise PermissionError(f"image {image!r} is not allowed on this node (allowed_images: {list(allowed)})")

    def _worker(self) -> None:
        while not self._stop.is_set():
            try:
                result = self.queue.get(timeout=0.5)
            except queue.Empty:
                continue
            threading.Thread(target=self._run_result, args=(result,), daemon=True).start()

    REPORT_BACKOFF_S = (0.5, 1, 2, 4, 8, 15, 30, 30, 30)      # a result outlives a server restart of about two minutes

    def _report_final(self, rid: int, final: dict) -> bool:
        """Hand a finished result to the server; a computed result is not thrown away because the server was unreachable
        for a moment (restart, network blip): retried with backoff.  What the server refuses for good (4xx: the result was
        deleted with its task, or is already finished) is not retried."""
        for attempt, pause in enumerate((0,) + self.REPORT_BACKOFF_S):
            if pause and self._stop.wait(pause):
                break
            try:
                self.client.request(f"result/{rid}", method="patch", json=final)
                if attempt:
                    log.info("result %s reported after %d retries", rid, attempt)
                return True
            except ServerError as e:
                if 400 <= e.status < 500 and e.status != 401:
                    log.error("the server refused result %s: %s", rid, e)
                    return False
                log.warning("could not report result %s (%s); retrying", rid, e)
            except Exception as e:  # noqa: BLE001 -- connection refused / reset / timeout
                log.warning("could not report result %s (%s: %s); retrying", rid, type(e).__name__, e)
        log.error("giving up on reporting result %s", rid)
        return False

    def _report_started(self, rid: int) -> None:
        try:
            self.client.request(f"result/{rid}", method="patch", json={"started_at": _now(), "status": "active"})
        except Exception as e:  # noqa: BLE001
            log.warning("could not report the start of result %s: %s", rid, e)

    def _run_result(self, result: dict) -> None:
        rid = result["id"]
        task = result["task"]
        log.info("starting task %s (result %s, image %s)", task["id"], rid, task.get("image"))
        marks = [("picked", time.perf_counter())]          # phase trace of the run (V6B200_TRACE_TASKS=1 logs it)

        def mark(name: str) -> None:
            marks.append((name, time.perf_counter()))
        started, start_reported = None, False
        logtxt, out_b64, status = "", None, "failed"
        try:
            self.check_image_allowed(task["image"])
            module = resolve_image(task["image"], self.config.get("algorithms"), bool(self.config.get("allow_module_images")))
            plain = self.cryptor.decrypt_str_to_bytes(result["input"]) if result.get("input") else b"{}"
            mark("resolve+decrypt")
            reply = self.client.request("token/container", method="post",
                                        json={"task_id": task["id"], "image": task["image"], "result_id": rid})
            token = reply["container_token"]
            mark("token")
            start_reported = True
            if not reply.get("started"):             # a server that does not take the start report with the token request
                started = threading.Thread(target=self._report_started, args=(rid,), daemon=True)
                started.start()
            run_dir = runtime_dir() / "volumes" / self.ctx.docker_temporary_volume_name(task["run_id"])
            work = run_dir / f"result-{rid}"
            work.mkdir(parents=True, exist_ok=True)
            (work / "input").write_bytes(plain)
            (work / "token").write_text(token)
            (work / "output").write_bytes(b"")
            mark("token+files")
            env = self._algorithm_env()
            label = (task.get("database") or "default")
            uri = self.ctx.databases.get(label) if hasattr(self.ctx, "databases") else None
            env.update({
                "INPUT_FILE": str(work / "input"), "OUTPUT_FILE": str(work / "output"), "TOKEN_FILE": str(work / "token"),
                "TEMPORARY_FOLDER": str(run_dir), "HOST": "http://127.0.0.1", "PORT": str(self.proxy.port), "API_PATH": "",
                "DATABASE_LABEL": label, "V6_ORGANIZATION_ID": str(self.client.organization_id),
                "V6_NODE_ID": str(self.client.node_id), "V6_COLLABORATION_ID": str(self.client.collaboration_id),
            })
            privacy = self.config.get("privacy") or {}             # the data station's floors for the tabular algorithms
            for key, var in (("min_rows", "V6B200_MIN_ROWS"), ("min_count", "V6B200_MIN_COUNT")):
                if privacy.get(key) is not None:
                    env[var] = str(int(privacy[key]))
            if uri is not None:
                env["DATABASE_URI"] = str(uri)
                env[f"{label.upper()}_DATABASE_URI"] = str(uri)
            if self.gpu is not None:
                env["V6_GPU"] = str(self.gpu)
            if self.gpu is not None and self.gpu_worker is not None and not self.gpu_worker_sock:
                # the resident worker is still coming up (first CUDA / torch import on a fresh box): a task that arrives now
                # waits for it instead of paying its own CUDA bring-up and leaving the worker cold for the next task
                deadline = time.time() + 150
                while (time.time() < deadline and not self.gpu_worker_sock and self.gpu_worker is not None
                       and self.gpu_worker.poll() is None and not self._stop.is_set()):
                    time.sleep(0.1)
            if self.gpu_worker_sock and self.gpu_worker is not None and self.gpu_worker.poll() is None:
                env["V6_GPU_WORKER"] = self.gpu_worker_sock
            pkg_root = str(Path(__file__).resolve().parent.parent.parent)
            env["PYTHONPATH"] = pkg_root + os.pathsep + env.get("PYTHONPATH", "")
            proc = self._launch_algorithm(module, env, work / "log")
            mark("spawned")
            self.running[rid] = proc
            self._task_of[rid] = task["id"]
            budget = float(self.config.get("task_timeout_s", 3600))
            if isinstance(proc, ZygoteProcess):
                proc.wait(timeout=budget)
                logtxt = proc.read_log()
            else:
                out, _ = proc.communicate(timeout=budget)
                logtxt = out.decode("utf-8", errors="replace")
            mark("algorithm done")
            if proc.returncode == 0:
                data = (work / "output").read_bytes()
                dest_org = task.get("initiator") or self.client.request(f"task/{task['id']}").get("initiator")
                out_b64 = self.encrypt_for_organization(data, int(dest_org)) if dest_org else self.cryptor.bytes_to_str(data)
                status = "completed"
            else:
                logtxt += f"\n[node] algorithm exited with code {proc.returncode}"
        except subprocess.TimeoutExpired:
            logtxt += "\n[node] algorithm timed out and was killed"
            self._kill_proc(self.running.get(rid))
        except Exception as e:  # noqa: BLE001
            log.exception("task %s failed", task.get("id"))
            logtxt += f"\n[node] failed to run algorithm: {e!r}"
        finally:
            self.running.pop(rid, None)
            self._task_of.pop(rid, None)
        if started is not None:
            started.join(timeout=30)
        final = {"finished_at": _now(), "result": out_b64, "log": logtxt[-20000:], "status": status}
        if not start_reported:                       # refused before it began (image policy, unknown image, bad input)
            final["started_at"] = final["finished_at"]
        self._report_final(rid, final)
        mark("reported")
        log.info("task %s result %s: %s", task.get("id"), rid, status)
        if os.environ.get("V6B200_TRACE_TASKS") == "1":
            log.info("result %s phases (ms): %s", rid, ", ".join(f"{b[0]} {1e3 * (b[1] - a[1]):.2f}" for a, b in zip(marks, marks[1:])))

    # environment variables an algorithm process inherits from the node (everything else is dropped: the reference
    # isolates algorithms in containers -- reference vantage6/cli/node.py:320 hands the node the docker socket for that --
    # here they are same-uid child processes, so at least the node's own secrets and unrelated settings stay out)
    ENV_PASSTHROUGH = ("PATH", "HOME", "LANG", "LC_ALL", "TMPDIR", "USER", "LD_LIBRARY_PATH", "PYTHONPATH", "VIRTUAL_ENV",
                       "CUDA_HOME", "CUDA_VISIBLE_DEVICES", "NVIDIA_VISIBLE_DEVICES", "XDG_RUNTIME_DIR", "V6B200_RUNTIME_DIR")
    ENV_PREFIXES = ("V6B200_", "NCCL_", "CUDA_", "TORCH_", "OMP_", "MKL_")
    ENV_BLOCKED = ("PRIVATE_KEY", "V6B200_ALLOW_PICKLE")

    def _algorithm_env(self) -> Dict[str, str]:
        env = {k: v for k, v in os.environ.items()
               if (k in self.ENV_PASSTHROUGH or k.startswith(self.ENV_PREFIXES)) and k not in self.ENV_BLOCKED}
        env["V6B200_ALLOW_PICKLE"] = "1" if self.config.get("allow_pickle") else "0"
        # directories (besides the per-run temporary folder and the log directory) that task inputs may name for checkpoints /
        # metrics: node config ``algorithm_data_dirs`` or the operator's V6_ALGORITHM_DATA_DIR (algorithm/builtin/fedavg.py::_confined)
        extra = list(self.config.get("algorithm_data_dirs") or []) + [d for d in os.environ.get("V6_ALGORITHM_DATA_DIR", "").split(os.pathsep) if d]
        if extra:
            env["V6_ALGORITHM_DATA_DIR"] = os.pathsep.join(str(d) for d in extra)
        log_dir = getattr(self.ctx, "log_dir", None)
        if log_dir:
            env["V6_LOG_DIR"] = str(log_dir)
        return env

    def _start_gpu_worker(self) -> None:
        """GPU-pinned nodes keep ONE resident process that owns the CUDA context, the NVLink symmetric heap and the
        trainers (model, optimizer state, captured CUDA graphs) across tasks; algorithm processes stay short-lived and
        talk to it over a Unix socket (node/gpu_worker.py).  ``V6B200_GPU_WORKER=0`` turns it off (one CUDA bring-up
        per task, the round-1 behaviour)."""
        if self.gpu is None or os.environ.get("V6B200_GPU_WORKER", "1") == "0":
            return
        sock = str(runtime_dir() / f"gpu-worker-{self.ctx.name}-{os.getpid()}.sock")
        env = self._algorithm_env()
        env["V6_GPU"] = str(self.gpu)
        pkg_root = str(Path(__file__).resolve().parent.parent.parent)
        env["PYTHONPATH"] = pkg_root + os.pathsep + env.get("PYTHONPATH", "")
        log_dir = getattr(self.ctx, "log_dir", None)
        out = open(Path(log_dir) / "gpu_worker.log", "ab") if log_dir else subprocess.DEVNULL
        self.gpu_worker = subprocess.Popen([sys.executable, "-m", "vantage6_b200.node.gpu_worker", "--socket", sock, "--gpu", str(self.gpu)],
                                           stdout=out, stderr=subprocess.STDOUT, env=env, start_new_session=True)
        deadline = time.time() + 180
        while time.time() < deadline and not self._stop.is_set():
            if self.gpu_worker.poll() is not None:
                log.error("GPU worker exited with code %s; tasks will bring up CUDA themselves", self.gpu_worker.returncode)
                self.gpu_worker = None
                return
            if os.path.exists(sock):
                try:
                    from .gpu_worker import call

                    call(sock, {"op": "ping"}, timeout=5)
                    self.gpu_worker_sock = sock
                    log.info("resident GPU worker up on GPU %s (pid %s)", self.gpu, self.gpu_worker.pid)
                    return
                except Exception:  # noqa: BLE001
                    pass
            time.sleep(0.2)
        log.warning("GPU worker did not come up in time")

    def _start_zygote(self) -> None:
        """Warm-start helper for the (short-lived) algorithm processes.  On GPU-pinned nodes the heavy state lives in the
        resident GPU worker, so their algorithm processes are as light as the CPU control-plane ones and fork from the
        zygote too; without the worker a GPU task needs a fresh interpreter (``V6B200_ZYGOTE=1`` forces the zygote on,
        ``=0`` off)."""
        default = "1" if (self.gpu is None or os.environ.get("V6B200_GPU_WORKER", "1") != "0") else "0"
        if os.environ.get("V6B200_ZYGOTE", default) == "0":
            return
        databases = getattr(self.ctx, "databases", None) or self.config.get("databases") or {}
        zygote = Zygote(runtime_dir(), preload=preload_for(list(databases.values()) if isinstance(databases, dict) else []))
        try:
            if zygote.start() and not self._stop.is_set():
                self.zygote = zygote
            else:
                zygote.stop()
        except Exception as e:  # noqa: BLE001
            log.warning("no warm-start helper (%s): one interpreter per task", e)

    def _launch_algorithm(self, module: str, env: Dict[str, str], log_path: Path):
        """Fork from the warm zygote when it is there (milliseconds), else start a fresh interpreter."""
        if self.zygote is not None and self.zygote.alive():
            try:
                return self.zygote.spawn(module, env, log_path)
            except Exception as e:  # noqa: BLE001
                log.warning("zygote spawn failed (%s); starting a fresh interpreter", e)
        return subprocess.Popen([sys.executable, "-m", "vantage6_b200.algorithm.wrapper", module],
                                stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env, start_new_session=True)

    @staticmethod
    def _kill_proc(proc: Optional[subprocess.Popen]) -> None:
        if proc is None:
            return
        try:
            os.killpg(os.getpgid(proc.pid), signal.SIGKILL)
        except Exception:  # noqa: BLE001
            try:
                proc.kill()
            except Exception:  # noqa: BLE001
                pass

    def kill_task(self, task_id=None) -> None:
        """Kill the algorithm processes of ``task_id`` (all of them when no task is named)."""
        for rid, proc in list(self.running.items()):
            if task_id is None or self._task_of.get(rid) == task_id:
                self._kill_proc(proc)

    # ------------------------------------------------------------------ lifecycle
    def start(self, block: bool = True) -> None:
        log.info("vantage6-b200 node runtime %s (gpu=%s)", __version__, self.gpu)
        self.authenticate()
        self.setup_encryption()
        self.proxy.start()
        threading.Thread(target=self._start_zygote, daemon=True).start()     # tasks arriving before it is up run cold
        threading.Thread(target=self._start_gpu_worker, daemon=True).start()
        self.sync_open_results()
        for target in (self._listen, self._worker, self._heartbeat):
            t = threading.Thread(target=target, daemon=True)
            t.start()
            self._threads.append(t)
        print(f"node '{self.client.name}' online (id={self.client.node_id}, gpu={self.gpu})", flush=True)
        if block:
            try:
                while not self._stop.wait(0.5):
                    pass
            except KeyboardInterrupt:
                pass
            self.stop()

    def stop(self) -> None:
        if self._stop.is_set() and not self._threads:
            return
        self._stop.set()
        for proc in list(self.running.values()):
            self._kill_proc(proc)
        try:
            self.client.request(f"node/{self.client.node_id}", method="patch", json={"status": "offline"})
        except Exception:  # noqa: BLE001
            pass
        self.proxy.stop()
        if self.gpu_worker is not None:
            try:
                from .gpu_worker import call

                call(self.gpu_worker_sock, {"op": "shutdown"}, timeout=5)
                self.gpu_worker.wait(timeout=10)
            except Exception:  # noqa: BLE001
                self._kill_proc(self.gpu_worker)
            self.gpu_worker = None
        if self.zygote is not None:
            self.zygote.stop()
            self.zygote = None
        self._threads = []
        log.info("node stopped")
