"""Small JSON-over-HTTP transport on the stdlib ``http.client`` with persistent connections.

The control plane of a federated round is a dozen tiny JSON requests between processes on one box (researcher ->
server, node -> server, algorithm -> node proxy -> server); with ``requests`` each of them costs ~0.6-1.0 ms of client-side
machinery (session / adapter / urllib3 pool / cookie jar), three to five times the server's own work.  This keeps a small
pool of keep-alive connections per host and does nothing else: measured 0.14 ms for ``GET /version`` and 0.41 ms for an
authenticated item against 0.63 / 0.99 ms (profiles/README.md, control plane section).

Environment proxies are never used (the peers are loopback or a configured server address).
"""
from __future__ import annotations

import http.client
import json as _json
import os
import select
import socket
import ssl
import threading
from typing import Any, Dict, Optional, Tuple
from urllib.parse import urlencode, urlsplit

_IDEMPOTENT = ("GET", "HEAD", "DELETE", "PUT")
_STALE = (ConnectionResetError, BrokenPipeError, ConnectionAbortedError, http.client.RemoteDisconnected,
          http.client.BadStatusLine, http.client.CannotSendRequest, http.client.ResponseNotReady,
          ssl.SSLEOFError, ssl.SSLZeroReturnError)          # a TLS peer that went away while the connection sat idle


class Response:
    __slots__ = ("status_code", "content")

    def __init__(self, status_code: int, content: bytes):
        self.status_code, self.content = status_code, content

    @property
    def text(self) -> str:
        return self.content.decode("utf-8", errors="replace")

    def json(self) -> Any:
        return _json.loads(self.content.decode("utf-8"))


class JsonHttp:
    """``request(method, url, json=, headers=, params=, timeout=) -> Response``; safe to share between threads: a
    connection is checked out of the idle pool for one request and put back when its response has been read, so
    short-lived threads (one per running task in a node) reuse warm connections instead of opening their own."""

    MAX_IDLE = 8           # per (scheme, host, port)

    def __init__(self, ca_file: Optional[str] = None):
        """``ca_file``: PEM bundle to verify an ``https`` server against (a private CA or the server's self-signed
        certificate); default: the system trust store.  ``$V6B200_CA_FILE`` supplies it where no argument can."""
        self._idle: Dict[Tuple[str, str, int], list] = {}
        self._lock = threading.Lock()
        self.ca_file = ca_file or os.environ.get("V6B200_CA_FILE") or None
        self._ssl_context = None

    def ssl_context(self):
        if self._ssl_context is None:
            self._ssl_context = ssl.create_default_context(cafile=self.ca_file)
        return self._ssl_context

    # ------------------------------------------------------------------ connections
    @staticmethod
    def _dropped(conn: http.client.HTTPConnection) -> bool:
        """An idle keep-alive connection that is readable has been closed by the peer (or holds garbage)."""
        sock = conn.sock
        if sock is None:
            return True
        if hasattr(sock, "pending"):        # TLS: post-handshake records (session tickets) make a healthy socket readable;
            return False                    # a dead one is caught by the retry in request()
        try:
            ready, _, _ = select.select([sock], [], [], 0)
            return bool(ready)
        except (OSError, ValueError):
            return True

    def _checkout(self, key: Tuple[str, str, int], timeout: float) -> Tuple[http.client.HTTPConnection, bool]:
        while True:
            with self._lock:
                idle = self._idle.get(key)
                conn = idle.pop() if idle else None
            if conn is None:
                break
            if not self._dropped(conn):
                conn.sock.settimeout(timeout)
                return conn, True
            conn.close()
        scheme, host, port = key
        if scheme == "https":
            conn = http.client.HTTPSConnection(host, port, timeout=timeout, context=self.ssl_context())
        else:
            conn = http.client.HTTPConnection(host, port, timeout=timeout)
        conn.connect()
        try:
            conn.sock.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
        except OSError:
            pass
        return conn, False

    def _checkin(self, key: Tuple[str, str, int], conn: http.client.HTTPConnection) -> None:
        with self._lock:
            idle = self._idle.setdefault(key, [])
            if len(idle) < self.MAX_IDLE:
                idle.append(conn)
                return
        conn.close()

    def idle_connections(self) -> int:
        with self._lock:
            return sum(len(v) for v in self._idle.values())

    def close(self) -> None:
        with self._lock:
            conns = [c for v in self._idle.values() for c in v]
            self._idle.clear()
        for conn in conns:
            try:
                conn.close()
            except Exception:  # noqa: BLE001
                pass

    # ------------------------------------------------------------------ requests
    def request(self, method: str, url: str, json: Any = None, headers: Optional[Dict[str, str]] = None,
                params: Optional[Dict[str, Any]] = None, timeout: float = 70.0) -> Response:
        method = method.upper()
        parts = urlsplit(url)
        scheme = parts.scheme or "http"
        key = (scheme, parts.hostname or "127.0.0.1", parts.port or (443 if scheme == "https" else 80))
        path = parts.path or "/"
        query = parts.query
        if params:
            extra = urlencode({k: v for k, v in params.items() if v is not None})
            query = f"{query}&{extra}" if query and extra else (query or extra)
        if query:
            path += "?" + query
        hdrs = {"Accept": "application/json", "Connection": "keep-alive"}
        body = None
        if json is not None:
            body = _json.dumps(json).encode("utf-8")
            hdrs["Content-Type"] = "application/json"
        elif method in ("POST", "PATCH", "PUT"):
            body = b""
        hdrs.update(headers or {})
        for attempt in (0, 1):
            conn, reused = self._checkout(key, timeout)
            try:
                conn.request(method, path, body=body, headers=hdrs)
                resp = conn.getresponse()
                data = resp.read()
                if resp.will_close:
                    conn.close()
                else:
                    self._checkin(key, conn)
                return Response(resp.status, data)
            except _STALE:
                # the peer closed a kept-alive connection between our check and our send: nothing was processed.  One
                # fresh attempt -- always for idempotent methods, for the others only if the connection was a reused one
                conn.close()
                if attempt == 0 and (method in _IDEMPOTENT or reused):
                    continue
                raise
            except BaseException:
                conn.close()
                raise
        raise RuntimeError("unreachable")
