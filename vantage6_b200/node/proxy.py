"""Node proxy server: the local HTTP endpoint through which a running algorithm (its
``ContainerClient``) reaches the central server (the reference names its host
``proxyserver``: reference vantage6/cli/globals.py:27; behaviour per SURVEY.md Appendix C).

It (1) forwards requests to the central server with the caller's container token,
(2) encrypts the inputs of sub-tasks per destination organization, and (3) decrypts results
addressed to this node's organization -- so algorithm code never sees keys or ciphertext.
"""
from __future__ import annotations

import json
import logging
import re
import threading
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from urllib.parse import urlsplit

from ..common import base64s_to_bytes
from ..common.jsonhttp import JsonHttp

log = logging.getLogger("proxy")


class _DaemonHTTPServer(ThreadingHTTPServer):
    daemon_threads = True          # request threads must not keep a stopping node alive


class ProxyServer:
    def __init__(self, node):
        self.node = node
        self._httpd = None
        self.port = None
        client_http = getattr(getattr(node, "client", None), "_http", None)
        self._http = client_http if isinstance(client_http, JsonHttp) else JsonHttp()      # same connections, same CA file

    def start(self) -> int:
        proxy = self

        class Handler(BaseHTTPRequestHandler):
            protocol_version = "HTTP/1.1"
            wbufsize = 64 * 1024                 # one send per response (no Nagle / delayed-ACK stall)
            disable_nagle_algorithm = True

            def log_message(self, fmt, *args):
                log.debug(fmt % args)

            def _serve(self, method):
                length = int(self.headers.get("Content-Length") or 0)
                raw = self.rfile.read(length) if length else b""
                body = json.loads(raw.decode("utf-8")) if raw else None
                status, payload = proxy.handle(method, self.path, body, self.headers.get("Authorization"))
                data = json.dumps(payload).encode("utf-8")
                self.send_response(status)
                self.send_header("Content-Type", "application/json")
                self.send_header("Content-Length", str(len(data)))
                self.end_headers()
                self.wfile.write(data)

            def do_GET(self):
                self._serve("GET")

            def do_POST(self):
                self._serve("POST")

            def do_PATCH(self):
                self._serve("PATCH")

            def do_DELETE(self):
                self._serve("DELETE")

        self._httpd = _DaemonHTTPServer(("127.0.0.1", 0), Handler)
        self.port = self._httpd.server_address[1]
        threading.Thread(target=self._httpd.serve_forever, kwargs={"poll_interval": 0.2}, daemon=True).start()
        log.info("proxy server listening on 127.0.0.1:%s", self.port)
        return self.port

    def stop(self) -> None:
        if self._httpd is not None:
            self._httpd.shutdown()
            self._httpd.server_close()
            self._httpd = None

    # ------------------------------------------------------------------
    def handle(self, method: str, path: str, body, auth):
        node = self.node
        parts = urlsplit(path)
        sub = parts.path
        url = node.client.generate_path_to(sub.lstrip("/")) + (("?" + parts.query) if parts.query else "")
        headers = {"Authorization": auth} if auth else {}
        try:
            if method == "POST" and re.fullmatch(r"/?task/?", sub) and isinstance(body, dict):
                # encrypt the (base64 plain) input for every destination organization
                for org in body.get("organizations", []):
                    plain = base64s_to_bytes(org["input"]) if isinstance(org.get("input"), str) else b""
                    org["input"] = node.encrypt_for_organization(plain, int(org["id"]))
            r = self._http.request(method, url, json=body, headers=headers, timeout=70)
            try:
                payload = r.json()
            except Exception:  # noqa: BLE001
                payload = {"msg": r.text}
            if r.status_code < 300 and method == "GET":
                if re.fullmatch(r"/?task/\d+/result/?", sub) and isinstance(payload, list):
                    for res in payload:
                        res["result"] = node.decrypt_to_plain_b64(res.get("result"))
                elif re.fullmatch(r"/?result/\d+/?", sub) and isinstance(payload, dict):
                    payload["result"] = node.decrypt_to_plain_b64(payload.get("result"))
                elif re.fullmatch(r"/?task/\d+/?", sub) and isinstance(payload, dict):
                    for res in payload.get("results") or []:               # ?include=results: a task with its results inline
                        if isinstance(res, dict) and "result" in res:
                            res["result"] = node.decrypt_to_plain_b64(res.get("result"))
            return r.status_code, payload
        except Exception as e:  # noqa: BLE001
            log.exception("proxy failure")
            return 502, {"msg": f"proxy error: {e}"}
