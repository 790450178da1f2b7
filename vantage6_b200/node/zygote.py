"""Warm start for algorithm runs.

vantage6 starts a fresh container per task; the process runtime started a fresh interpreter per task, and
on this box ~0.45 s of a 0.5 s two-node task was the interpreter importing numpy and the wrapper.  The
*zygote* is a small single-threaded helper process the node starts once: it imports the algorithm wrapper (and
the data-loading libraries) up front and then ``fork()``s one child per task.  A child is a process of its own
(own session / process group, own environment, own working files -- the isolation a task gets is the same as
before), it just starts in milliseconds.  CUDA is never touched in the zygote, so children are free to
initialise it.

Protocol (Unix stream socket, one connection per task):

    node   -> zygote   {"module": "...", "env": {...}, "log": "/path/to/log"}\\n
    zygote -> node     {"pid": 12345}\\n                  (right after the fork)
    zygote -> node     {"exit": 0}\\n                     (when that child has exited; then the connection closes)

``V6B200_ZYGOTE=0`` makes the node fall back to one interpreter per task.
"""
from __future__ import annotations

import json
import os
import select
import signal
import socket
import subprocess
import sys
import time
from pathlib import Path
from typing import Dict, Optional


# ------------------------------------------------------------------------------------------ server side
def _run_child(request: dict) -> None:
    """In the forked child: become the algorithm process described by ``request`` and never return."""
    code = 1
    try:
        os.setsid()                                            # own process group: the node kills with killpg
        log_fd = os.open(request["log"], os.O_WRONLY | os.O_CREAT | os.O_APPEND, 0o600)
        os.dup2(log_fd, 1)
        os.dup2(log_fd, 2)
        os.close(log_fd)
        devnull = os.open(os.devnull, os.O_RDONLY)
        os.dup2(devnull, 0)
        os.close(devnull)
        env = {str(k): str(v) for k, v in request["env"].items()}
        os.environ.clear()
        os.environ.update(env)
        for entry in reversed([p for p in env.get("PYTHONPATH", "").split(os.pathsep) if p]):
            if entry not in sys.path:
                sys.path.insert(0, entry)
        for sig in (signal.SIGTERM, signal.SIGINT, signal.SIGCHLD):
            signal.signal(sig, signal.SIG_DFL)
        from ..algorithm.wrapper import run_algorithm

        sys.argv = ["vantage6_b200.algorithm.wrapper", request["module"]]
        code = int(run_algorithm(request["module"]) or 0)
    except SystemExit as e:
        code = e.code if isinstance(e.code, int) else (0 if e.code is None else 1)
    except BaseException:  # noqa: BLE001 -- anything escaping an algorithm ends the run, with a trace in its log
        import traceback

        traceback.print_exc()
        code = 1
    finally:
        try:
            sys.stdout.flush()
            sys.stderr.flush()
        finally:
            os._exit(code & 0xFF)


PRELOAD = ("numpy", "jwt", "vantage6_b200.client", "vantage6_b200.common.jsonhttp",
           "vantage6_b200.algorithm.builtin.average", "vantage6_b200.algorithm.builtin.weighted_mean")


def preload_for(database_uris) -> tuple:
    """Modules worth importing once in the zygote for a node with these databases.  pandas only where a tabular file
    will be read: with it (and the thread pools and shared objects it drags in) every fork costs ~4 ms instead of ~1 ms,
    which a node that serves ``.npy`` / ``.pt`` / synthetic data should not pay on each task."""
    tabular = any(str(u).lower().rsplit(".", 1)[-1] in ("csv", "parquet", "xlsx", "tsv") for u in (database_uris or []))
    return PRELOAD + (("pandas",) if tabular else ())


def serve(socket_path: str, preload=None) -> None:
    """Zygote main loop: accept task descriptions, fork, report exits.  Single-threaded on purpose."""
    from ..algorithm import wrapper  # noqa: F401  (the point of the zygote: pay for these imports once)

    for optional in (preload if preload is not None else PRELOAD + ("pandas",)):
        try:
            __import__(optional)
        except Exception:  # noqa: BLE001
            pass
    parent = os.getppid()
    try:
        os.unlink(socket_path)
    except FileNotFoundError:
        pass
    listener = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    listener.bind(socket_path)
    os.chmod(socket_path, 0o600)
    listener.listen(64)
    children: Dict[int, socket.socket] = {}
    # a child's exit must wake the loop at once: SIGCHLD writes a byte into this pipe
    wake_r, wake_w = os.pipe()
    os.set_blocking(wake_r, False)
    os.set_blocking(wake_w, False)
    signal.signal(signal.SIGCHLD, lambda *_: None)
    signal.set_wakeup_fd(wake_w, warn_on_full_buffer=False)
    print("zygote ready", flush=True)
    while True:
        ready, _, _ = select.select([listener, wake_r], [], [], 0.25)
        if wake_r in ready:
            try:
                os.read(wake_r, 4096)
            except OSError:
                pass
        if listener in ready:
            conn, _ = listener.accept()
            try:
                request = json.loads(conn.makefile("r").readline())
                pid = os.fork()
                if pid == 0:
                    signal.set_wakeup_fd(-1)
                    os.close(wake_r)
                    os.close(wake_w)
                    listener.close()
                    for other in children.values():
                        other.close()
                    conn.close()
                    _run_child(request)                         # never returns
                children[pid] = conn
                conn.sendall((json.dumps({"pid": pid}) + "\n").encode())
            except Exception as e:  # noqa: BLE001
                try:
                    conn.sendall((json.dumps({"error": repr(e)}) + "\n").encode())
                finally:
                    conn.close()
        while children:                                          # reap whatever has finished
            try:
                pid, status = os.waitpid(-1, os.WNOHANG)
            except ChildProcessError:
                break
            if pid == 0:
                break
            conn = children.pop(pid, None)
            if conn is not None:
                code = os.waitstatus_to_exitcode(status)
                try:
                    conn.sendall((json.dumps({"exit": code}) + "\n").encode())
                except OSError:
                    pass
                conn.close()
        if os.getppid() != parent and not children:              # the node is gone: nothing left to serve
            return


# ------------------------------------------------------------------------------------------ client side
class ZygoteProcess:
    """Handle of one forked algorithm process (the part of ``subprocess.Popen`` the node uses)."""

    def __init__(self, pid: int, conn: socket.socket, log_path: Path):
        self.pid = pid
        self._conn = conn
        self._buf = b""
        self.log_path = log_path
        self.returncode: Optional[int] = None

    def _readline(self, timeout: Optional[float]) -> str:
        """One protocol line; raises ``subprocess.TimeoutExpired`` (and can be called again afterwards)."""
        deadline = None if timeout is None else time.time() + timeout
        while b"\n" not in self._buf:
            left = None if deadline is None else deadline - time.time()
            if left is not None and left <= 0:
                raise subprocess.TimeoutExpired(cmd="algorithm", timeout=timeout)
            ready, _, _ = select.select([self._conn], [], [], left)
            if not ready:
                raise subprocess.TimeoutExpired(cmd="algorithm", timeout=timeout)
            chunk = self._conn.recv(4096)
            if not chunk:                                   # the zygote went away
                break
            self._buf += chunk
        line, _, self._buf = self._buf.partition(b"\n")
        return line.decode("utf-8", errors="replace")

    def wait(self, timeout: Optional[float] = None) -> int:
        if self.returncode is None:
            line = self._readline(timeout)
            reply = json.loads(line) if line else {"exit": -9}
            self.returncode = int(reply.get("exit", -9))
            self._conn.close()
        return self.returncode

    def kill(self) -> None:
        try:
            os.killpg(self.pid, signal.SIGKILL)
        except (ProcessLookupError, PermissionError):
            pass

    def read_log(self) -> str:
        try:
            return Path(self.log_path).read_text(errors="replace")
        except OSError:
            return ""


class Zygote:
    """Owns the helper process; ``spawn`` asks it for a child."""

    def __init__(self, work_dir: Path, preload=None):
        self.preload = preload
        self.socket_path = str(Path(work_dir) / f"zygote-{os.getpid()}.sock")
        if len(self.socket_path) > 100:                           # AF_UNIX path limit
            import tempfile

            self.socket_path = str(Path(tempfile.gettempdir()) / f"v6zyg-{os.getpid()}.sock")
        self._proc: Optional[subprocess.Popen] = None

    def start(self, timeout: float = 30.0) -> bool:
        env = dict(os.environ)
        pkg_root = str(Path(__file__).resolve().parent.parent.parent)
        env["PYTHONPATH"] = pkg_root + os.pathsep + env.get("PYTHONPATH", "")
        argv = [sys.executable, "-m", "vantage6_b200.node.zygote", self.socket_path]
        if self.preload is not None:
            argv.append(",".join(self.preload))
        self._proc = subprocess.Popen(argv, env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, stdin=subprocess.DEVNULL)
        deadline = time.time() + timeout
        line = b""
        while time.time() < deadline and self._proc.poll() is None:
            ready, _, _ = select.select([self._proc.stdout], [], [], 0.2)
            if ready:
                line = self._proc.stdout.readline()
                break
        return b"zygote ready" in line

    def alive(self) -> bool:
        return self._proc is not None and self._proc.poll() is None

    def spawn(self, module: str, env: Dict[str, str], log_path: Path) -> ZygoteProcess:
        conn = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        conn.settimeout(10.0)
        conn.connect(self.socket_path)
        conn.sendall((json.dumps({"module": module, "env": env, "log": str(log_path)}) + "\n").encode())
        conn.settimeout(None)
        handle = ZygoteProcess(0, conn, log_path)
        reply = json.loads(handle._readline(10.0) or "{}")
        if "pid" not in reply:
            conn.close()
            raise RuntimeError(f"zygote refused the task: {reply.get('error', 'no reply')}")
        handle.pid = int(reply["pid"])
        return handle

    def stop(self) -> None:
        if self._proc is not None:
            try:
                self._proc.terminate()
                self._proc.wait(timeout=3)
            except Exception:  # noqa: BLE001
                try:
                    self._proc.kill()
                except Exception:  # noqa: BLE001
                    pass
            self._proc = None
        try:
            os.unlink(self.socket_path)
        except OSError:
            pass


if __name__ == "__main__":
    serve(sys.argv[1], tuple(m for m in sys.argv[2].split(",") if m) if len(sys.argv) > 2 else None)
