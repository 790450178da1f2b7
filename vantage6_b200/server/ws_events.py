"""WebSocket event channel of the central server.

The reference serves its event channel as websockets under ``uwsgi --http :5000 --gevent 1000 --http-websockets``
(reference vantage6/cli/server.py:224): a thousand listeners cost a thousand greenlets, not a thousand threads.  The REST
side of this server runs on a threading HTTP server where a long-poll listener pins one thread; this module is the
scalable path: ONE asyncio thread (``websockets``) serves every listener, events are pushed the moment ``EventBus.emit``
is called, and the long-poll endpoint (``GET /event``) stays as the fallback for clients without the package.

Protocol: ``ws://host:event_port/?token=<JWT access token>&since=<last event id>[&task_id=<id>]``; the server first
replays the buffered events after ``since`` that the identity may see (same rooms as ``GET /event``), then pushes one JSON
text frame ``{"id", "name", "data"}`` per event.  The port is advertised by ``GET /health`` (``event_port``).
"""
from __future__ import annotations

import asyncio
import json
import logging
import threading
from typing import List, Optional
from urllib.parse import parse_qs, urlsplit

log = logging.getLogger("server.ws")


class WebSocketEvents:
    def __init__(self, app, ip: str = "127.0.0.1", port: int = 0, ssl_context=None):
        self.app, self.ip, self.port, self.ssl_context = app, ip, port, ssl_context
        self.loop: Optional[asyncio.AbstractEventLoop] = None
        self._thread: Optional[threading.Thread] = None
        self._server = None
        self._subs: List[tuple] = []                # (queue, rooms)
        self._ready = threading.Event()
        self.connections = 0

    # ------------------------------------------------------------------ bus side (any thread)
    def publish(self, ev: dict) -> None:
        loop = self.loop
        if loop is None or not self._subs:
            return
        loop.call_soon_threadsafe(self._fanout, ev)

    def _fanout(self, ev: dict) -> None:
        rooms = set(ev.get("rooms", []))
        msg = {k: ev[k] for k in ("id", "name", "data")}
        for q, sub_rooms in list(self._subs):
            if rooms & sub_rooms:
                q.put_nowait(msg)

    # ------------------------------------------------------------------ connection handler
    async def _handler(self, ws) -> None:
        path = getattr(getattr(ws, "request", None), "path", None) or getattr(ws, "path", "/")
        q = parse_qs(urlsplit(path).query)
        try:
            ident = self.app.identity({"Authorization": "Bearer " + (q.get("token") or [""])[0]})
            ident = self.app.require(ident)
            rooms = set(self.app.event_rooms(ident, (q.get("task_id") or [None])[0]))
        except Exception as e:  # noqa: BLE001 -- HTTPError and friends: refuse the subscription
            await ws.close(code=4401, reason=str(getattr(e, "msg", e))[:100])
            return
        since = int((q.get("since") or [self.app.events.last_id()])[0])
        queue: asyncio.Queue = asyncio.Queue()
        sub = (queue, rooms)
        self._subs.append(sub)                      # subscribe first, then replay: nothing falls between the two
        self.connections += 1
        try:
            sent = since
            for ev in self.app.events.wait(since, list(rooms), 0.0):
                await ws.send(json.dumps(ev))
                sent = max(sent, ev["id"])
            while True:
                ev = await queue.get()
                if ev["id"] > sent:
                    await ws.send(json.dumps(ev))
                    sent = ev["id"]
        except Exception:  # noqa: BLE001 -- closed connection
            pass
        finally:
            self.connections -= 1
            if sub in self._subs:
                self._subs.remove(sub)

    # ------------------------------------------------------------------ lifecycle
    def start(self) -> int:
        import websockets

        def run():
            self.loop = asyncio.new_event_loop()
            asyncio.set_event_loop(self.loop)

            async def boot():
                self._server = await websockets.serve(self._handler, self.ip, self.port, ping_interval=20, max_size=1 << 20,
                                                      ssl=self.ssl_context)
                self.port = self._server.sockets[0].getsockname()[1]
                self._ready.set()
            self.loop.run_until_complete(boot())
            self.loop.run_forever()

        self._thread = threading.Thread(target=run, daemon=True, name="ws-events")
        self._thread.start()
        if not self._ready.wait(10):
            raise RuntimeError("websocket event channel did not start")
        log.info("event channel (websocket) on ws://%s:%s", self.ip, self.port)
        return self.port

    def stop(self) -> None:
        loop = self.loop
        if loop is None:
            return

        async def shut():
            if self._server is not None:
                self._server.close()
                await self._server.wait_closed()
            loop.stop()
        asyncio.run_coroutine_threadsafe(shut(), loop)
        if self._thread is not None:
            self._thread.join(timeout=5)
        self.loop = None
