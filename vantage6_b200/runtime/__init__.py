"""Process runtime -- the Docker-daemon replacement.

The reference CLI manages *containers* through the docker SDK (``docker.from_env()``,
``containers.list(filters={"label": ...})``, ``containers.run(image, command=..., volumes=...,
labels=..., environment=..., name=..., detach=True)``, ``containers.get(name).stop()/kill()/
attach()/exec_run()``, ``volumes.create/list``; reference vantage6/cli/node.py:240-410,
server.py:130-248).  On an 8xB200 box a federated node is not a container but ONE PROCESS PINNED
TO ONE GPU, so this module exposes the same object model over plain processes:

* a "container" is a detached child process; its registry entry (pid, labels, name, command,
  log file, environment) is a JSON file under ``<runtime_dir>/containers/`` -- the registry plays
  the role of docker labels for ``list``/``stop``/``attach``;
* an "image" names a python entry point (``vnode-local``, ``vserver-local`` ... are resolved to
  ``python -m vantage6_b200.cli.*``); reference image names are accepted and mapped;
* "volumes" are directories under ``<runtime_dir>/volumes/``; mount tables (``/mnt/...`` targets)
  are resolved to host paths inside the command line and exported as ``V6_MOUNTS``;
* ``attach`` tails the process's log file; ``exec_run`` runs a command in the same environment.

GPU pinning: ``environment={"V6_GPU": "k"}`` (set by ``vnode start --gpu k``); every GPU stays visible so that
peer memory can be mapped (parallel/symm.py).
"""
from __future__ import annotations

import json
import os
import shlex
import shutil
import signal
import subprocess
import sys
import tempfile
import time
from pathlib import Path
from typing import Dict, Iterator, List, Optional

from ..common.globals import APPNAME, HOME_ENV


class APIError(Exception):
    """Raised when the runtime refuses an operation (docker.errors.APIError equivalent)."""


class NotFound(APIError):
    pass


class errors:  # namespace shim so call sites read like the docker SDK: runtime.errors.APIError
    APIError = APIError
    NotFound = NotFound


def runtime_dir() -> Path:
    home = os.environ.get(HOME_ENV)
    if home:
        return Path(home) / "runtime"
    return Path(os.environ.get("XDG_RUNTIME_DIR", tempfile.gettempdir())) / f"{APPNAME}-runtime-{os.getuid()}"


def _pid_alive(pid: int) -> bool:
    if pid <= 0:
        return False
    try:
        os.kill(pid, 0)
    except ProcessLookupError:
        return False
    except PermissionError:
        return True
    # zombie check
    try:
        with open(f"/proc/{pid}/stat") as f:
            return f.read().split(") ")[-1].split()[0] != "Z"
    except OSError:
        return True


# entry points that reference command strings name
ENTRY_POINTS = {
    "vnode-local": [sys.executable, "-m", "vantage6_b200.cli.node_local"],
    "vserver-local": [sys.executable, "-m", "vantage6_b200.cli.server_local"],
    "uwsgi": [sys.executable, "-m", "vantage6_b200.cli.server_local", "uwsgi"],
    "v6-mq-broker": [sys.executable, "-m", "vantage6_b200.server.mq_broker"],
    "rabbitmqctl": [sys.executable, "-m", "vantage6_b200.server.mq_broker", "ctl"],
}


class ExecResult:
    def __init__(self, exit_code: int, output: bytes):
        self.exit_code = exit_code
        self.output = output


class Container:
    def __init__(self, runtime: "LocalRuntime", meta: dict):
        self._rt = runtime
        self.meta = meta

    # docker-SDK-like attributes
    @property
    def name(self) -> str:
        return self.meta["name"]

    @property
    def id(self) -> str:
        return f"{self.meta['name']}:{self.meta['pid']}"

    @property
    def labels(self) -> Dict[str, str]:
        return self.meta.get("labels", {})

    @property
    def pid(self) -> int:
        return int(self.meta["pid"])

    @property
    def status(self) -> str:
        return "running" if _pid_alive(self.pid) else "exited"

    def __repr__(self):
        return f"<Container: {self.id}>"

    def _signal(self, sig) -> None:
        try:
            os.killpg(os.getpgid(self.pid), sig)
        except (ProcessLookupError, PermissionError):
            try:
                os.kill(self.pid, sig)
            except ProcessLookupError:
                pass

    def kill(self) -> None:
        self._signal(signal.SIGKILL)
        self._wait_gone(5)
        self._rt._unregister(self.name)

    def stop(self, timeout: float = 10) -> None:
        """SIGTERM, then SIGKILL after ``timeout`` seconds (docker's 10 s grace:
        reference vantage6/cli/node.py:461-463)."""
        self._signal(signal.SIGTERM)
        if not self._wait_gone(timeout):
            self._signal(signal.SIGKILL)
            self._wait_gone(5)
        self._rt._unregister(self.name)

    def remove(self, force: bool = False) -> None:
        if _pid_alive(self.pid):
            if not force:
                raise APIError(f"container {self.name} is running")
            self.kill()
        self._rt._unregister(self.name)

    def _wait_gone(self, timeout: float) -> bool:
        t0 = time.time()
        while time.time() - t0 < timeout:
            if not _pid_alive(self.pid):
                return True
            try:        # reap if it is our child
                os.waitpid(self.pid, os.WNOHANG)
            except ChildProcessError:
                pass
            time.sleep(0.05)
        return not _pid_alive(self.pid)

    def logs(self, stream: bool = False, stdout: bool = True, follow: Optional[bool] = None, **_):
        path = Path(self.meta["log_file"])
        if not stream:
            return path.read_bytes() if path.exists() else b""
        return self._tail(path, follow=True if follow is None else follow)

    def attach(self, stream: bool = True, logs: bool = True, stdout: bool = True, **_):
        return self._tail(Path(self.meta["log_file"]), follow=True, from_start=logs)

    def _tail(self, path: Path, follow: bool, from_start: bool = True) -> Iterator[bytes]:
        def gen():
            while not path.exists():
                if not _pid_alive(self.pid):
                    return
                time.sleep(0.1)
            with open(path, "rb") as f:
                if not from_start:
                    f.seek(0, os.SEEK_END)
                while True:
                    line = f.readline()
                    if line:
                        yield line
                        continue
                    if not follow or not _pid_alive(self.pid):
                        rest = f.read()
                        if rest:
                            yield rest
                        return
                    time.sleep(0.2)
        return gen()

    def exec_run(self, cmd, stdout: bool = True, **_) -> ExecResult:
        argv = self._rt._resolve_command(cmd, self.meta.get("mounts", {}))
        env = dict(os.environ)
        env.update(self.meta.get("environment") or {})
        try:
            p = subprocess.run(argv, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env=env, timeout=120)
            return ExecResult(p.returncode, p.stdout)
        except Exception as e:  # noqa: BLE001
            return ExecResult(1, str(e).encode())

    def wait(self, timeout: Optional[float] = None) -> dict:
        t0 = time.time()
        while _pid_alive(self.pid):
            try:
                os.waitpid(self.pid, os.WNOHANG)
            except ChildProcessError:
                pass
            if timeout is not None and time.time() - t0 > timeout:
                raise TimeoutError(self.name)
            time.sleep(0.05)
        return {"StatusCode": 0}


class ContainerCollection:
    def __init__(self, runtime: "LocalRuntime"):
        self._rt = runtime

    def list(self, filters: Optional[dict] = None, all: bool = False) -> List[Container]:  # noqa: A002
        out = []
        for meta in self._rt._entries():
            c = Container(self._rt, meta)
            alive = _pid_alive(c.pid)
            if not alive and not all:
                if not meta.get("keep", False):
                    self._rt._unregister(c.name)         # auto_remove semantics
                continue
            if filters and "label" in filters:
                wanted = filters["label"]
                wanted = [wanted] if isinstance(wanted, str) else wanted
                ok = True
                for w in wanted:
                    k, _, v = w.partition("=")
                    if k not in c.labels or (v and str(c.labels[k]) != v):
                        ok = False
                if not ok:
                    continue
            if filters and "name" in filters and filters["name"] not in c.name:
                continue
            out.append(c)
        return out

    def get(self, name: str) -> Container:
        meta = self._rt._entry(name)
        if meta is None:
            raise NotFound(f"No such container: {name}")
        return Container(self._rt, meta)

    def run(self, image: str, command=None, volumes=None, mounts=None, detach: bool = True, labels=None,
            environment=None, name: Optional[str] = None, auto_remove: bool = True, tty: bool = False, ports=None,
            network=None, restart_policy=None, hostname=None, **_) -> Container:
        mount_table: Dict[str, str] = {}
        if isinstance(volumes, dict):                       # {source: {"bind": target, "mode": ..}}
            for src, spec in volumes.items():
                mount_table[str(spec["bind"])] = str(src)
        elif volumes:                                       # ["source:target", ...]
            for v in volumes:
                src, _, tgt = str(v).rpartition(":")
                mount_table[tgt] = self._rt.volume_path(src)
        for m in mounts or []:                              # Mount(target, source, type="bind")
            mount_table[str(m.target)] = str(m.source)
        argv = self._rt._resolve_command(command if command is not None else image, mount_table)
        name = name or f"{APPNAME}-anon-{int(time.time() * 1000)}"
        if self._rt._entry(name) is not None and _pid_alive(int(self._rt._entry(name)["pid"])):
            raise APIError(f"Conflict. The container name {name!r} is already in use")
        log_dir = self._rt.root / "logs"
        log_dir.mkdir(parents=True, exist_ok=True)
        log_file = log_dir / f"{name}.log"
        env = dict(os.environ)
        env.update({k: str(v) for k, v in (environment or {}).items()})
        env["V6_MOUNTS"] = json.dumps(mount_table)
        env["V6_CONTAINER_NAME"] = name
        if ports:
            env["V6_PORTS"] = json.dumps({k: list(v) if isinstance(v, (tuple, list)) else v for k, v in ports.items()})
        pkg_root = str(Path(__file__).resolve().parent.parent.parent)
        env["PYTHONPATH"] = pkg_root + os.pathsep + env.get("PYTHONPATH", "")
        with open(log_file, "ab") as lf:
            proc = subprocess.Popen(argv, stdout=lf, stderr=subprocess.STDOUT, stdin=subprocess.DEVNULL, env=env,
                                    start_new_session=True)
        meta = {"name": name, "pid": proc.pid, "image": image, "command": argv, "labels": dict(labels or {}),
                "environment": {k: str(v) for k, v in (environment or {}).items()}, "mounts": mount_table,
                "log_file": str(log_file), "keep": not auto_remove, "started_at": time.time(),
                "network": network, "restart_policy": restart_policy}
        self._rt._register(meta)
        c = Container(self._rt, meta)
        c._proc = proc
        if not detach:
            proc.wait()
        return c


class Volume:
    def __init__(self, runtime: "LocalRuntime", name: str):
        self._rt = runtime
        self.name = name

    @property
    def path(self) -> Path:
        return self._rt.root / "volumes" / self.name

    def remove(self) -> None:
        try:
            shutil.rmtree(self.path)
        except FileNotFoundError:
            raise NotFound(self.name)
        except OSError as e:
            raise APIError(str(e))


class VolumeCollection:
    def __init__(self, runtime: "LocalRuntime"):
        self._rt = runtime

    def create(self, name: str, **_) -> Volume:
        v = Volume(self._rt, name)
        v.path.mkdir(parents=True, exist_ok=True)
        return v

    def list(self) -> List[Volume]:
        d = self._rt.root / "volumes"
        if not d.exists():
            return []
        return [Volume(self._rt, p.name) for p in sorted(d.iterdir()) if p.is_dir()]

    def get(self, name: str) -> Volume:
        v = Volume(self._rt, name)
        if not v.path.exists():
            raise NotFound(name)
        return v


class ImageCollection:
    """Images are python entry points shipped with the package: pulling is a no-op that
    succeeds for known names and raises for unknown ones (best-effort, like the reference's
    pull: reference vantage6/cli/node.py:297-305)."""

    def pull(self, image: str, **_):
        return image

    def get(self, image: str):
        return image


class Mount:
    """docker.types.Mount look-alike (target, source, type)."""

    def __init__(self, target: str, source: str, type: str = "bind", **_):  # noqa: A002
        self.target, self.source, self.type = target, source, type


class types:  # namespace shim: runtime.types.Mount
    Mount = Mount


class LocalRuntime:
    def __init__(self, root: Optional[Path] = None):
        self.root = Path(root) if root else runtime_dir()
        self._containers = ContainerCollection(self)
        self._volumes = VolumeCollection(self)
        self._images = ImageCollection()

    @property
    def images(self) -> ImageCollection:
        return self._images

    # class-level properties so tests can patch ``LocalRuntime.containers`` / ``.volumes`` the
    # way the reference tests patch ``docker.DockerClient.containers``
    @property
    def containers(self) -> ContainerCollection:
        return self._containers

    @property
    def volumes(self) -> VolumeCollection:
        return self._volumes

    # -- docker.from_env() / ping ------------------------------------------------------------
    def ping(self) -> bool:
        (self.root / "containers").mkdir(parents=True, exist_ok=True)
        if not os.access(self.root, os.W_OK):
            raise APIError(f"runtime directory {self.root} is not writable")
        return True

    def volume_path(self, name_or_path: str) -> str:
        if os.path.isabs(name_or_path) or name_or_path.startswith("."):
            return name_or_path
        return str(self.root / "volumes" / name_or_path)

    # -- registry ----------------------------------------------------------------------------
    def _reg_file(self, name: str) -> Path:
        return self.root / "containers" / f"{name}.json"

    def _register(self, meta: dict) -> None:
        f = self._reg_file(meta["name"])
        f.parent.mkdir(parents=True, exist_ok=True)
        tmp = f.with_suffix(".tmp")
        tmp.write_text(json.dumps(meta))
        tmp.replace(f)

    def _unregister(self, name: str) -> None:
        try:
            self._reg_file(name).unlink()
        except FileNotFoundError:
            pass

    def _entry(self, name: str) -> Optional[dict]:
        f = self._reg_file(name)
        if not f.exists():
            return None
        try:
            return json.loads(f.read_text())
        except Exception:  # noqa: BLE001
            return None

    def _entries(self) -> List[dict]:
        d = self.root / "containers"
        if not d.exists():
            return []
        out = []
        for f in sorted(d.glob("*.json")):
            try:
                out.append(json.loads(f.read_text()))
            except Exception:  # noqa: BLE001
                continue
        return out

    # -- command resolution ------------------------------------------------------------------
    def _resolve_command(self, command, mount_table: Dict[str, str]) -> List[str]:
        argv = shlex.split(command) if isinstance(command, str) else list(command)
        if not argv:
            raise APIError("empty command")
        head = ENTRY_POINTS.get(argv[0])
        if head is not None:
            argv = head + argv[1:]
        # translate container paths (/mnt/...) to host paths using the mount table
        targets = sorted(mount_table, key=len, reverse=True)
        out = []
        for a in argv:
            for t in targets:
                tt = t.rstrip("/")
                if a == tt or a.startswith(tt + "/"):
                    a = mount_table[t].rstrip("/") + a[len(tt):]
                    break
            out.append(a)
        return out


def from_env() -> LocalRuntime:
    return LocalRuntime()
