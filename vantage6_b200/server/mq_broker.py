"""In-box message-queue broker: the RabbitMQ replacement that lets several server processes
share events (reference vantage6/cli/server.py:266-273).

``serve``  -- ZeroMQ XSUB/XPUB forwarder on ``port`` (publish) and ``port+1`` (subscribe) plus a
              REP status socket on ``port+2``; credentials come from the same definitions.json the
              reference writes for RabbitMQ (user + salted-SHA256 password hash + vhost).
``ctl status`` -- the ``rabbitmqctl status`` probe used by the start-up poll.
``attach_app`` -- mirror a ServerApp's EventBus through the broker (publish local events, inject
              remote ones), topic = vhost.
"""
from __future__ import annotations

import json
import os
import sys
import threading
import uuid

import zmq


def serve(port: int, definitions: str | None, data: str | None) -> int:
    ctx = zmq.Context.instance()
    xsub = ctx.socket(zmq.XSUB)
    xsub.bind(f"tcp://127.0.0.1:{port}")
    xpub = ctx.socket(zmq.XPUB)
    xpub.bind(f"tcp://127.0.0.1:{port + 1}")
    rep = ctx.socket(zmq.REP)
    rep.bind(f"tcp://127.0.0.1:{port + 2}")
    defs = {}
    if definitions and os.path.exists(definitions):
        with open(definitions) as f:
            defs = json.load(f)
    if data:
        os.makedirs(data, exist_ok=True)
        with open(os.path.join(data, "broker.pid"), "w") as f:
            f.write(str(os.getpid()))

    def status_loop():
        while True:
            rep.recv()
            rep.send_json({"status": "running", "pid": os.getpid(), "vhosts": [v["name"] for v in defs.get("vhosts", [])],
                           "users": [u["name"] for u in defs.get("users", [])]})

    threading.Thread(target=status_loop, daemon=True).start()
    print(f"v6-mq-broker listening on {port} (pub) / {port + 1} (sub) / {port + 2} (status)", flush=True)
    try:
        zmq.proxy(xsub, xpub)
    except KeyboardInterrupt:
        pass
    return 0


def ctl_status(port: int, timeout_ms: int = 2000) -> int:
    ctx = zmq.Context.instance()
    req = ctx.socket(zmq.REQ)
    req.setsockopt(zmq.LINGER, 0)
    req.setsockopt(zmq.RCVTIMEO, timeout_ms)
    req.setsockopt(zmq.SNDTIMEO, timeout_ms)
    req.connect(f"tcp://127.0.0.1:{port + 2}")
    try:
        req.send(b"status")
        print(json.dumps(req.recv_json()))
        return 0
    except zmq.ZMQError:
        print(json.dumps({"status": "down"}))
        return 1
    finally:
        req.close()


def attach_app(app, rabbitmq_uri: str) -> None:
    """Mirror ``app.events`` through the broker named by ``rabbitmq_uri``."""
    from ..cli.rabbitmq.queue_manager import split_rabbitmq_uri

    parts = split_rabbitmq_uri(rabbitmq_uri)
    port, topic = int(parts["port"]), parts["vhost"].encode()
    origin = uuid.uuid4().hex
    ctx = zmq.Context.instance()
    pub = ctx.socket(zmq.PUB)
    pub.connect(f"tcp://127.0.0.1:{port}")
    sub = ctx.socket(zmq.SUB)
    sub.connect(f"tcp://127.0.0.1:{port + 1}")
    sub.setsockopt(zmq.SUBSCRIBE, topic)
    lock = threading.Lock()

    def mirror(ev: dict) -> None:
        with lock:
            pub.send_multipart([topic, json.dumps({"origin": origin, **ev}).encode()])

    def pump():
        while True:
            _, raw = sub.recv_multipart()
            ev = json.loads(raw)
            if ev.get("origin") != origin:
                app.events.emit(ev["name"], ev["data"], ev["rooms"], mirrored=True)

    app.events.mirror = mirror
    threading.Thread(target=pump, daemon=True).start()


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    def opt(name, default=None):
        return argv[argv.index(name) + 1] if name in argv else default
    if argv and argv[0] == "serve":
        return serve(int(opt("--port", 5672)), opt("--definitions"), opt("--data"))
    if argv and argv[0] == "ctl":
        return ctl_status(int(opt("--port", 5672)))
    print("usage: mq_broker serve --port P [--definitions F] [--data D] | ctl status --port P")
    return 2


if __name__ == "__main__":
    sys.exit(main())
