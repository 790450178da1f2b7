"""Central server: REST API + event channel (the ``vantage6-server`` runtime that the reference
CLI launches with ``uwsgi --http :5000 --gevent 1000 --http-websockets ... wsgi.py``:
reference vantage6/cli/server.py:223-228; resources per SURVEY.md Appendix C).

* JSON REST resources under ``api_path``: ``/token/{user,node,container,refresh}``,
  ``/organization``, ``/collaboration``, ``/node``, ``/user``, ``/role``, ``/rule``, ``/task``,
  ``/result``, ``/health``, ``/version``, ``/event``; role management, ``/recover/{lost,reset}``,
  ``/password/change`` and the membership sub-resources live in ``admin_routes.py``.
* JWT identities of three kinds (user, node, container) signed with ``jwt_secret_key``.
* Rule-based permissions: rule = (resource, scope in {own, organization, collaboration, global},
  operation in {view, create, edit, delete}); default roles are created on first start.
* Event channel: nodes long-poll ``GET /event?since=<id>`` (rooms per collaboration / node) --
  the push semantics of the reference's Socket.IO namespace without a websocket stack.  With
  ``rabbitmq_uri`` configured, events are mirrored through the message-queue sidecar so that
  several server processes can share them (reference server.py:266-273).

Runs on the stdlib ``ThreadingHTTPServer``: on an 8-GPU box the control plane is a few small
JSON messages per round; tensor bytes never travel on it (parallel/symm.py carries those).
"""
from __future__ import annotations

import collections
import datetime as _dt
import json
import logging
import os
import re
import socket
import threading
import time
import traceback
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer
from typing import Any, Callable, Dict, List, Optional, Tuple
from urllib.parse import parse_qs, urlsplit

import jwt

from .._version import __version__
from . import admin_routes
from .db import Database, check_password, hash_password, now

log = logging.getLogger("server")
TRACE_HTTP = os.environ.get("V6B200_TRACE_HTTP") == "1"        # log every request with its handling time


class _ApiHTTPServer(ThreadingHTTPServer):
    daemon_threads = True          # request threads must not keep a stopping server alive
    request_queue_size = 128       # 8 nodes x (heartbeat + long-poll + proxy traffic) connect in bursts

SCOPES = ["own", "organization", "collaboration", "global"]
OPERATIONS = ["view", "create", "edit", "delete"]
RESOURCES = ["user", "organization", "collaboration", "role", "node", "task", "result", "port", "event"]

DEFAULT_ROLES = {
    "Root": ("Super role", [(r, "global", o) for r in RESOURCES for o in OPERATIONS]),
    "Collaboration Admin": ("Can manage a collaboration", [(r, "collaboration", o) for r in ("organization", "collaboration", "node", "task", "result", "user", "port") for o in OPERATIONS]),
    "Organization Admin": ("Can manage an organization", [(r, "organization", o) for r in ("user", "organization", "node", "task", "result", "role", "port") for o in OPERATIONS] + [("collaboration", "organization", "view")]),
    "Researcher": ("Can create tasks and view results", [("task", "organization", "view"), ("task", "organization", "create"), ("result", "organization", "view"), ("organization", "collaboration", "view"), ("collaboration", "organization", "view"), ("node", "organization", "view"), ("user", "organization", "view"), ("port", "organization", "view")]),
    "Viewer": ("Can view tasks and results", [("task", "organization", "view"), ("result", "organization", "view"), ("organization", "organization", "view"), ("collaboration", "organization", "view"), ("node", "organization", "view")]),
}


MAX_BODY_BYTES = 256 << 20          # task inputs / results are small; model tensors never travel on this API


class PlainText(str):
    """A handler's reply that goes out as ``text/plain`` instead of JSON (``GET /metrics``)."""


class HTTPError(Exception):
    def __init__(self, status: int, msg: str):
        super().__init__(msg)
        self.status, self.msg = status, msg


class EventBus:
    """In-memory event log with blocking reads (long-poll)."""

    HEAVY_KEPT = 256

    def __init__(self, maxlen: int = 10000):
        self._events: collections.deque = collections.deque(maxlen=maxlen)
        self._heavy: collections.deque = collections.deque()       # recent events that carry a work item (see emit)
        self._cond = threading.Condition()
        self._next_id = 1
        self.mirror: Optional[Callable[[dict], None]] = None
        self.push: Optional[Callable[[dict], None]] = None     # websocket fan-out (server/ws_events.py)

    def emit(self, name: str, data: dict, rooms: List[str], mirrored: bool = False) -> int:
        with self._cond:
            ev = {"id": self._next_id, "name": name, "data": data, "rooms": rooms, "ts": time.time()}
            self._next_id += 1
            self._events.append(ev)
            if "result" in data:
                # a work item rides on its event so that the node can start at once; the log keeps thousands of events, the
                # payloads only of the most recent ones (a listener that far behind fetches the item by id instead)
                self._heavy.append(ev)
                if len(self._heavy) > self.HEAVY_KEPT:
                    old = self._heavy.popleft()
                    old["data"] = {k: v for k, v in old["data"].items() if k != "result"}
            self._cond.notify_all()
        if self.push is not None:
            try:
                self.push(ev)
            except Exception:  # noqa: BLE001
                log.debug("event push failed", exc_info=True)
        if self.mirror is not None and not mirrored:
            try:
                self.mirror({"name": name, "data": data, "rooms": rooms})
            except Exception:  # noqa: BLE001
                log.debug("event mirror failed", exc_info=True)
        return ev["id"]

    def last_id(self) -> int:
        with self._cond:
            return self._next_id - 1

    def wait(self, since: int, rooms: List[str], timeout: float) -> List[dict]:
        deadline = time.time() + timeout
        with self._cond:
            while True:
                out = [e for e in self._events if e["id"] > since and (set(e["rooms"]) & set(rooms))]
                if out:
                    return [{k: e[k] for k in ("id", "name", "data")} for e in out]
                left = deadline - time.time()
                if left <= 0:
                    return []
                self._cond.wait(min(left, 1.0))


class ServerApp:
    """The WSGI-``app`` equivalent: owns the database, the event bus and the route table."""

    def __init__(self, config: dict, db: Optional[Database] = None, data_dir=None, name: str = "server"):
        self.config = config
        self.name = name
        self.api_path = (config.get("api_path") or "/api").rstrip("/")
        self.db = db or Database(config.get("uri", "sqlite://"), data_dir, bool(config.get("allow_drop_all", False)))
        self.secret = config.get("jwt_secret_key") or self.db.token_secret(None)
        self.events = EventBus()
        self.started_at = time.time()
        self.token_expiry_s = int(config.get("token_expires_hours", 6) * 3600)
        self._routes: List[Tuple[str, re.Pattern, Callable]] = []
        self._conns: set = set()                                       # open keep-alive connections (closed by stop())
        self._conns_lock = threading.Lock()
        self._stats: Dict[Tuple[str, str, int], List[float]] = {}      # (method, route, status) -> [count, seconds]
        self._stats_lock = threading.Lock()
        self._httpd: Optional[ThreadingHTTPServer] = None
        self.ws = None                                # websocket event channel (server/ws_events.py)
        self._thread: Optional[threading.Thread] = None
        self._register_routes()
        self.ensure_defaults()

    # ------------------------------------------------------------------ bootstrap
    def ensure_defaults(self) -> None:
        db = self.db
        for res in RESOURCES:
            for sc in SCOPES:
                for op in OPERATIONS:
                    db.execute("INSERT OR IGNORE INTO rule (name, operation, scope, description) VALUES (?,?,?,?)",
                               (res, op, sc, f"{op} {res} ({sc})"))
        for rname, (desc, rules) in DEFAULT_ROLES.items():
            if db.one("SELECT id FROM role WHERE name=? AND organization_id IS NULL", (rname,)) is None:
                rid = db.insert("role", name=rname, description=desc, organization_id=None)
                for (res, sc, op) in rules:
                    rule = db.one("SELECT id FROM rule WHERE name=? AND scope=? AND operation=?", (res, sc, op))
                    db.execute("INSERT OR IGNORE INTO role_rule VALUES (?,?)", (rid, rule["id"]))
        if db.one("SELECT id FROM user LIMIT 1") is None:
            org = db.one("SELECT id FROM organization WHERE name='root'")
            oid = org["id"] if org else db.insert("organization", name="root")
            uid = db.insert("user", username="root", password=hash_password("root"), firstname="root",
                            lastname="root", email="root@localhost", organization_id=oid)
            root_role = db.one("SELECT id FROM role WHERE name='Root'")
            db.execute("INSERT OR IGNORE INTO user_role VALUES (?,?)", (uid, root_role["id"]))
            log.warning("Created default root user (username 'root', password 'root') -- change the password!")

    # ------------------------------------------------------------------ tokens
    def make_token(self, kind: str, ident: dict, refresh: bool = False) -> str:
        exp = _dt.datetime.now(_dt.timezone.utc) + _dt.timedelta(seconds=self.token_expiry_s * (8 if refresh else 1))
        payload = {"sub": json.dumps({"type": kind, **ident}), "exp": exp, "typ": "refresh" if refresh else "access"}
        return jwt.encode(payload, self.secret, algorithm="HS256")

    def identity(self, headers) -> Optional[dict]:
        auth = headers.get("Authorization", "")
        if not auth.startswith("Bearer "):
            return None
        try:
            payload = jwt.decode(auth[7:], self.secret, algorithms=["HS256"])
        except jwt.ExpiredSignatureError:
            raise HTTPError(401, "Token has expired")
        except jwt.PyJWTError:
            raise HTTPError(401, "Invalid token")
        ident = json.loads(payload["sub"])
        ident["_typ"] = payload.get("typ", "access")
        return ident

    # ------------------------------------------------------------------ permissions
    def scope_of(self, ident: dict, resource: str, operation: str) -> Optional[str]:
        """Widest scope the identity holds for (resource, operation)."""
        if ident["type"] != "user":
            return None
        best = -1
        for r in self.db.user_rules(ident["id"]):
            if r["name"] == resource and r["operation"] == operation:
                best = max(best, SCOPES.index(r["scope"]))
        return SCOPES[best] if best >= 0 else None

    def _orgs_in_reach(self, ident: dict, scope: Optional[str]) -> Optional[set]:
        """Organization ids visible at ``scope`` (None = all)."""
        if scope == "global":
            return None
        org = ident.get("organization_id")
        if scope == "collaboration":
            orgs = {org}
            for c in self.db.organization_collaborations(org):
                orgs |= set(self.db.collaboration_organizations(c))
            return orgs
        if scope in ("organization", "own"):
            return {org}
        return set()

    def require(self, ident: Optional[dict], *kinds: str) -> dict:
        if ident is None:
            raise HTTPError(401, "Missing Authorization Header")
        if ident.get("_typ") == "refresh":
            raise HTTPError(401, "Only access tokens are allowed")
        if kinds and ident["type"] not in kinds:
            raise HTTPError(403, f"Not allowed for identity type {ident['type']!r}")
        # a token outlives nothing: the account (or the node / the task of a container token) must still be there
        table, key = {"user": ("user", "id"), "node": ("node", "id"), "container": ("task", "task_id")}[ident["type"]]
        if self.db.one(f"SELECT id FROM {table} WHERE id=?", (ident.get(key),)) is None:
            raise HTTPError(401, f"This {ident['type']} token belongs to a deleted {table}")
        return ident

    def can_view_collaboration(self, ident: dict, collaboration_id: int, resource: str = "task") -> bool:
        """May ``ident`` see items (tasks, results, nodes, members) of this collaboration?  Nodes and containers only
        their own collaboration; users need a view rule on ``resource``: global scope sees everything, any narrower
        scope only collaborations their organization takes part in (the same reach the list endpoints filter by)."""
        if ident["type"] in ("node", "container"):
            return ident.get("collaboration_id") == collaboration_id
        sc = self.scope_of(ident, resource, "view")
        if sc == "global":
            return True
        return sc is not None and collaboration_id in set(self.db.organization_collaborations(ident["organization_id"]))

    def event_rooms(self, ident: dict, task_id=None) -> List[str]:
        """Event rooms an identity may listen to (shared by the long-poll endpoint and the websocket channel)."""
        if ident["type"] == "user":
            rooms = [f"collaboration_{c}" for c in self.db.organization_collaborations(ident["organization_id"])]
            if self.scope_of(ident, "event", "view") == "global":
                rooms = [f"collaboration_{c['id']}" for c in self.db.query("SELECT id FROM collaboration")]
        else:
            rooms = [f"collaboration_{ident['collaboration_id']}"]
            if ident["type"] == "node":
                rooms.append(f"node_{ident['id']}")
        if task_id is not None:
            t = self.db.get("task", int(task_id))
            if t is None:
                raise HTTPError(404, f"task id={task_id} is not found")
            self.require_collaboration_view(ident, t["collaboration_id"], "task")
            rooms.append(f"task_{t['id']}")
        return rooms

    def require_collaboration_view(self, ident: dict, collaboration_id: int, resource: str = "task") -> None:
        if not self.can_view_collaboration(ident, collaboration_id, resource):
            raise HTTPError(401, "You lack the permission to do that!")

    def check_grant(self, ident: dict, role_ids, rule_ids, edit_scope: Optional[str]) -> None:
        """A user may only hand out permissions it holds itself: every (resource, operation, scope) implied by the
        requested roles / rules must be covered by one of the caller's own rules at an equal or wider scope, and a
        caller without global user-edit scope may only assign default roles or roles of its own organization."""
        implied = []
        for rid in role_ids:
            role = self.db.get("role", int(rid))
            if role is None:
                raise HTTPError(404, f"role id={rid} not found")
            if edit_scope != "global" and role.get("organization_id") not in (None, ident["organization_id"]):
                raise HTTPError(401, f"You cannot assign role id={rid} of another organization")
            implied += self.db.query("SELECT rule.* FROM rule JOIN role_rule ON rule.id = role_rule.rule_id WHERE role_rule.role_id=?",
                                     (int(rid),))
        for rid in rule_ids:
            rule = self.db.get("rule", int(rid))
            if rule is None:
                raise HTTPError(404, f"rule id={rid} not found")
            implied.append(rule)
        for r in implied:
            mine = self.scope_of(ident, r["name"], r["operation"])
            if mine is None or SCOPES.index(mine) < SCOPES.index(r["scope"]):
                raise HTTPError(401, f"You cannot grant {r['name']}/{r['operation']} at scope {r['scope']!r}: "
                                     "you do not hold that permission yourself")

    # ------------------------------------------------------------------ serialisation
    def link(self, resource: str, id_: int) -> dict:
        return {"id": id_, "link": f"{self.api_path}/{resource}/{id_}", "methods": ["GET", "PATCH", "DELETE"]}

    def org_json(self, o: dict) -> dict:
        db = self.db
        return {**{k: o.get(k) for k in ("id", "name", "domain", "address1", "address2", "zipcode", "country", "public_key")},
                "collaborations": [self.link("collaboration", c) for c in db.organization_collaborations(o["id"])],
                "users": [self.link("user", u["id"]) for u in db.query("SELECT id FROM user WHERE organization_id=?", (o["id"],))],
                "nodes": [self.link("node", n["id"]) for n in db.query("SELECT id FROM node WHERE organization_id=?", (o["id"],))]}

    def collab_json(self, c: dict) -> dict:
        db = self.db
        return {"id": c["id"], "name": c["name"], "encrypted": bool(c["encrypted"]),
                "organizations": [self.link("organization", o) for o in db.collaboration_organizations(c["id"])],
                "nodes": [self.link("node", n["id"]) for n in db.query("SELECT id FROM node WHERE collaboration_id=?", (c["id"],))],
                "tasks": [self.link("task", t["id"]) for t in db.query("SELECT id FROM task WHERE collaboration_id=?", (c["id"],))]}

    def node_json(self, n: dict, with_key: bool = False) -> dict:
        out = {"id": n["id"], "name": n["name"], "ip": n["ip"], "status": n["status"], "last_seen": n["last_seen"],
               "gpu": n.get("gpu"), "type": "node",
               "collaboration": self.link("collaboration", n["collaboration_id"]) if n["collaboration_id"] is not None else None,
               "organization": self.link("organization", n["organization_id"])}
        if with_key:
            out["api_key"] = n["api_key"]
        return out

    def user_json(self, u: dict) -> dict:
        db = self.db
        return {"id": u["id"], "username": u["username"], "firstname": u["firstname"], "lastname": u["lastname"],
                "email": u["email"], "type": "user", "last_seen": u["last_seen"],
                "organization": self.link("organization", u["organization_id"]),
                "roles": [self.link("role", r["role_id"]) for r in db.query("SELECT role_id FROM user_role WHERE user_id=?", (u["id"],))],
                "rules": [self.link("rule", r["rule_id"]) for r in db.query("SELECT rule_id FROM user_rule WHERE user_id=?", (u["id"],))]}

    def result_json(self, r: dict, with_task: bool = False) -> dict:
        out = {"id": r["id"], "input": r["input"], "result": r["result"], "log": r["log"],
               "assigned_at": r["assigned_at"], "started_at": r["started_at"], "finished_at": r["finished_at"],
               "status": r["status"], "organization": self.link("organization", r["organization_id"]),
               "task": self.link("task", r["task_id"])}
        if with_task:
            t = self.db.get("task", r["task_id"])
            out["task"] = {**self.link("task", t["id"]), "name": t["name"], "image": t["image"], "database": t["database"],
                           "run_id": t["run_id"], "collaboration_id": t["collaboration_id"], "parent_id": t["parent_id"],
                           "initiator": t["initiator_id"]}
        return out

    def task_json(self, t: dict, include_results: bool = False) -> dict:
        db = self.db
        results = db.query("SELECT * FROM result WHERE task_id=? ORDER BY id", (t["id"],))
        return {"id": t["id"], "name": t["name"], "description": t["description"], "image": t["image"],
                "database": t["database"], "run_id": t["run_id"], "created_at": t["created_at"],
                "initiator": t["initiator_id"], "init_user": t["init_user_id"],
                "collaboration": self.link("collaboration", t["collaboration_id"]),
                "parent": self.link("task", t["parent_id"]) if t["parent_id"] else None,
                "children": [self.link("task", c["id"]) for c in db.query("SELECT id FROM task WHERE parent_id=?", (t["id"],))],
                "complete": all(r["finished_at"] is not None for r in results),
                "results": [self.result_json(r) if include_results else self.link("result", r["id"]) for r in results]}

    # ------------------------------------------------------------------ routing
    def route(self, method: str, pattern: str):
        def deco(fn):
            self._routes.append((method, re.compile("^" + pattern + "/?$"), fn))
            return fn
        return deco

    def dispatch(self, method: str, path: str, query: Dict[str, List[str]], body: Any, headers) -> Tuple[int, Any]:
        if not path.startswith(self.api_path):
            return 404, {"msg": f"unknown path {path}"}
        sub = path[len(self.api_path):] or "/"
        t0, route = time.perf_counter(), "(unmatched)"
        status, payload = 404, {"msg": f"no route for {method} {sub}"}
        try:
            ident = self.identity(headers)
            for m, rx, fn in self._routes:
                if m != method:
                    continue
                mt = rx.match(sub)
                if mt:
                    route = rx.pattern[1:-3].replace("(\\d+)", "<id>")
                    q = {k: v[0] for k, v in query.items()}
                    res = fn(ident, body if isinstance(body, dict) else {}, q, *mt.groups())
                    status, payload = (res[1], res[0]) if isinstance(res, tuple) else (200, res)
                    break
        except HTTPError as e:
            status, payload = e.status, {"msg": e.msg}
        except Exception as e:  # noqa: BLE001
            log.error("unhandled error in %s %s: %s", method, path, traceback.format_exc())
            status, payload = 500, {"msg": f"internal server error: {e}"}
        if route != "/event":                   # long polls would only measure their own timeout
            with self._stats_lock:
                st = self._stats.setdefault((method, route, status), [0, 0.0])
                st[0] += 1
                st[1] += time.perf_counter() - t0
        return status, payload

    # ------------------------------------------------------------------ resources
    def _register_routes(self) -> None:  # noqa: C901 -- a flat route table reads best in one place
        app, db = self, self.db

        # ---- meta
        @app.route("GET", "/version")
        def version(ident, body, q):
            return {"version": __version__}

        @app.route("GET", "/health")
        def health(ident, body, q):
            db.one("SELECT 1 AS ok")
            return {"database": True, "uptime_s": time.time() - app.started_at, "events": app.events.last_id(),
                    "event_port": app.ws.port if app.ws is not None else None,
                    "event_listeners": app.ws.connections if app.ws is not None else 0}

        # ---- tokens
        @app.route("POST", "/token/user")
        def token_user(ident, body, q):
            username, password = body.get("username"), body.get("password")
            if not username or not password:
                raise HTTPError(400, "Username and/or password missing in JSON body")
            u = db.one("SELECT * FROM user WHERE username=?", (username,))
            policy = app.config.get("password_policy") or {}
            max_failed, lock_min = int(policy.get("max_failed_attempts", 5)), float(policy.get("inactivation_minutes", 15))
            if u is not None and (u["failed_login_attempts"] or 0) >= max_failed and u.get("last_login_attempt"):
                since = (_dt.datetime.now(_dt.timezone.utc) - _dt.datetime.fromisoformat(u["last_login_attempt"])).total_seconds()
                if since < lock_min * 60:
                    raise HTTPError(401, f"Your account is blocked for the next {max(1, int(lock_min - since / 60))} minutes due to failed "
                                         "login attempts. Please wait or reactivate your account via the password recovery.")
                db.update("user", u["id"], failed_login_attempts=0)
                u["failed_login_attempts"] = 0
            if u is None or not check_password(password, u["password"]):
                if u is not None:
                    db.update("user", u["id"], failed_login_attempts=(u["failed_login_attempts"] or 0) + 1, last_login_attempt=now())
                raise HTTPError(401, "Invalid username or password!")
            db.update("user", u["id"], last_seen=now(), failed_login_attempts=0)
            ident_ = {"id": u["id"], "organization_id": u["organization_id"]}
            return {"access_token": app.make_token("user", ident_), "refresh_token": app.make_token("user", ident_, True),
                    "user_url": f"{app.api_path}/user/{u['id']}", "refresh_url": f"{app.api_path}/token/refresh"}

        @app.route("POST", "/token/node")
        def token_node(ident, body, q):
            key = body.get("api_key")
            if not key:
                raise HTTPError(400, "api_key missing in JSON body")
            n = db.one("SELECT * FROM node WHERE api_key=?", (key,))
            if n is None:
                raise HTTPError(401, "Api key is not recognized!")
            if n["collaboration_id"] is None:
                raise HTTPError(401, "This node is not attached to a collaboration")
            db.update("node", n["id"], last_seen=now(), status="online", ip=body.get("ip"), gpu=body.get("gpu"))
            ident_ = {"id": n["id"], "organization_id": n["organization_id"], "collaboration_id": n["collaboration_id"]}
            app.events.emit("node-status-changed", {"id": n["id"], "name": n["name"], "online": True},
                            [f"collaboration_{n['collaboration_id']}"])
            return {"access_token": app.make_token("node", ident_), "refresh_token": app.make_token("node", ident_, True),
                    "node_url": f"{app.api_path}/node/{n['id']}", "refresh_url": f"{app.api_path}/token/refresh"}

        @app.route("POST", "/token/container")
        def token_container(ident, body, q):
            ident = app.require(ident, "node")
            task_id, image = body.get("task_id"), body.get("image")
            t = db.get("task", int(task_id)) if task_id else None
            if t is None:
                raise HTTPError(404, f"Task {task_id} does not exist")
            if t["collaboration_id"] != ident["collaboration_id"]:
                raise HTTPError(401, "Task does not belong to the node's collaboration")
            if t["image"] != image:
                raise HTTPError(401, "Node is not allowed to issue a token for a different image")
            if db.task_complete(t["id"]):
                raise HTTPError(400, "Task is already finished")
            ident_ = {"node_id": ident["id"], "organization_id": ident["organization_id"],
                      "collaboration_id": ident["collaboration_id"], "task_id": t["id"], "image": image}
            reply = {"container_token": app.make_token("container", ident_)}
            if body.get("result_id") is not None:
                # asking for the algorithm's token is the first thing a node does for a work item: with ``result_id`` the
                # same request reports the start (one round trip less than a separate PATCH /result/<id>)
                r = db.get("result", int(body["result_id"]))
                if r is not None and r["task_id"] == t["id"] and r["organization_id"] == ident["organization_id"] and r["finished_at"] is None:
                    if r["started_at"] is None:
                        db.update("result", r["id"], started_at=now(), status="active")
                        app.events.emit("status_update", {"result_id": r["id"], "task_id": t["id"], "status": "active",
                                                          "organization_id": r["organization_id"], "parent_id": t["parent_id"],
                                                          "task_complete": False},
                                        [f"collaboration_{t['collaboration_id']}", f"task_{t['id']}"])
                    reply["started"] = True
            return reply

        @app.route("POST", "/token/refresh")
        def token_refresh(ident, body, q):
            if ident is None or ident.get("_typ") != "refresh":
                raise HTTPError(401, "A refresh token is required")
            kind = ident.pop("type")
            ident.pop("_typ")
            return {"access_token": app.make_token(kind, ident)}

        # ---- organization
        @app.route("GET", "/organization")
        def org_list(ident, body, q):
            ident = app.require(ident)
            rows = db.query("SELECT * FROM organization ORDER BY id")
            if ident["type"] == "user":
                reach = app._orgs_in_reach(ident, app.scope_of(ident, "organization", "view"))
            else:
                reach = set(db.collaboration_organizations(ident["collaboration_id"]))
            return [app.org_json(o) for o in rows if reach is None or o["id"] in reach]

        @app.route("POST", "/organization")
        def org_create(ident, body, q):
            ident = app.require(ident, "user")
            if app.scope_of(ident, "organization", "create") != "global":
                raise HTTPError(401, "You lack the permission to do that!")
            if not body.get("name"):
                raise HTTPError(400, "name is required")
            if db.one("SELECT id FROM organization WHERE name=?", (body["name"],)):
                raise HTTPError(400, f"Organization {body['name']!r} already exists")
            oid = db.insert("organization", **{k: body.get(k) for k in ("name", "domain", "address1", "address2", "zipcode", "country", "public_key")})
            return app.org_json(db.get("organization", oid)), 201

        @app.route("GET", r"/organization/(\d+)")
        def org_get(ident, body, q, oid):
            ident = app.require(ident)
            o = db.get("organization", int(oid))
            if o is None:
                raise HTTPError(404, f"Organization id={oid} not found")
            if ident["type"] == "user":
                reach = app._orgs_in_reach(ident, app.scope_of(ident, "organization", "view"))
                if reach is not None and o["id"] not in reach and o["id"] != ident["organization_id"]:
                    raise HTTPError(401, "You do not have permission to view this organization")
            elif o["id"] not in db.collaboration_organizations(ident["collaboration_id"]):
                raise HTTPError(401, "You do not have permission to view this organization")
            return app.org_json(o)

        @app.route("PATCH", r"/organization/(\d+)")
        def org_patch(ident, body, q, oid):
            """``PATCH /organization/<id> {"public_key": b64}`` is what ``vnode create-private-key``
            calls (reference vantage6/cli/node.py:610-614)."""
            ident = app.require(ident, "user", "node")
            o = db.get("organization", int(oid))
            if o is None:
                raise HTTPError(404, f"Organization id={oid} not found")
            if ident["type"] == "user":
                sc = app.scope_of(ident, "organization", "edit")
                if not (sc == "global" or (sc in ("organization", "collaboration") and o["id"] == ident["organization_id"])):
                    raise HTTPError(401, "You do not have permission to edit this organization")
            elif ident["organization_id"] != o["id"]:
                raise HTTPError(401, "A node can only edit its own organization")
            editable = ("public_key",) if ident["type"] == "node" else ("name", "domain", "address1", "address2", "zipcode", "country", "public_key")
            fields = {k: body[k] for k in editable if k in body}       # a node publishes its organization's key, nothing else
            if ident["type"] == "node" and set(body) - set(editable):
                raise HTTPError(401, "A node may only update its organization's public key")
            db.update("organization", o["id"], **fields)
            return app.org_json(db.get("organization", o["id"]))

        # ---- collaboration
        @app.route("GET", "/collaboration")
        def collab_list(ident, body, q):
            ident = app.require(ident)
            rows = db.query("SELECT * FROM collaboration ORDER BY id")
            if ident["type"] == "user":
                sc = app.scope_of(ident, "collaboration", "view")
                if sc != "global":
                    mine = set(db.organization_collaborations(ident["organization_id"])) if sc else set()
                    rows = [c for c in rows if c["id"] in mine]
            else:
                rows = [c for c in rows if c["id"] == ident["collaboration_id"]]
            return [app.collab_json(c) for c in rows]

        @app.route("POST", "/collaboration")
        def collab_create(ident, body, q):
            ident = app.require(ident, "user")
            if app.scope_of(ident, "collaboration", "create") != "global":
                raise HTTPError(401, "You lack the permission to do that!")
            if not body.get("name"):
                raise HTTPError(400, "name is required")
            if db.one("SELECT id FROM collaboration WHERE name=?", (body["name"],)):
                raise HTTPError(400, f"Collaboration {body['name']!r} already exists")
            cid = db.insert("collaboration", name=body["name"], encrypted=1 if body.get("encrypted") else 0)
            for oid in body.get("organization_ids", []):
                if db.get("organization", int(oid)) is None:
                    raise HTTPError(400, f"Organization id={oid} does not exist")
                db.execute("INSERT OR IGNORE INTO member VALUES (?,?)", (cid, int(oid)))
            return app.collab_json(db.get("collaboration", cid)), 201

        @app.route("GET", r"/collaboration/(\d+)")
        def collab_get(ident, body, q, cid):
            ident = app.require(ident)
            c = db.get("collaboration", int(cid))
            if c is None:
                raise HTTPError(404, f"collaboration id={cid} can not be found")
            if ident["type"] != "user":
                if ident["collaboration_id"] != c["id"]:
                    raise HTTPError(401, "not your collaboration")
            elif app.scope_of(ident, "collaboration", "view") != "global" and \
                    c["id"] not in db.organization_collaborations(ident["organization_id"]):
                raise HTTPError(401, "You do not have permission to view this collaboration")
            return app.collab_json(c)

        @app.route("PATCH", r"/collaboration/(\d+)")
        def collab_patch(ident, body, q, cid):
            ident = app.require(ident, "user")
            c = db.get("collaboration", int(cid))
            if c is None:
                raise HTTPError(404, f"collaboration id={cid} can not be found")
            sc = app.scope_of(ident, "collaboration", "edit")
            if not (sc == "global" or (sc == "collaboration" and c["id"] in db.organization_collaborations(ident["organization_id"]))):
                raise HTTPError(401, "You lack the permission to do that!")
            if "name" in body:
                db.update("collaboration", c["id"], name=body["name"])
            if "encrypted" in body:
                db.update("collaboration", c["id"], encrypted=1 if body["encrypted"] else 0)
            if "organization_ids" in body:
                db.execute("DELETE FROM member WHERE collaboration_id=?", (c["id"],))
                for oid in body["organization_ids"]:
                    db.execute("INSERT OR IGNORE INTO member VALUES (?,?)", (c["id"], int(oid)))
            return app.collab_json(db.get("collaboration", c["id"]))

        @app.route("DELETE", r"/collaboration/(\d+)")
        def collab_delete(ident, body, q, cid):
            ident = app.require(ident, "user")
            if app.scope_of(ident, "collaboration", "delete") != "global":
                raise HTTPError(401, "You lack the permission to do that!")
            db.execute("DELETE FROM member WHERE collaboration_id=?", (int(cid),))
            db.delete("collaboration", int(cid))
            return {"msg": f"collaboration id={cid} successfully deleted"}

        @app.route("GET", r"/collaboration/(\d+)/organization")
        def collab_orgs(ident, body, q, cid):
            app.require_collaboration_view(app.require(ident), int(cid), "collaboration")
            return [app.org_json(db.get("organization", o)) for o in db.collaboration_organizations(int(cid))]

        @app.route("GET", r"/collaboration/(\d+)/node")
        def collab_nodes(ident, body, q, cid):
            app.require_collaboration_view(app.require(ident), int(cid), "node")
            return [app.node_json(n) for n in db.query("SELECT * FROM node WHERE collaboration_id=?", (int(cid),))]

        @app.route("GET", r"/collaboration/(\d+)/task")
        def collab_tasks(ident, body, q, cid):
            app.require_collaboration_view(app.require(ident), int(cid), "task")
            return [app.task_json(t) for t in db.query("SELECT * FROM task WHERE collaboration_id=? ORDER BY id", (int(cid),))]

        # ---- node
        @app.route("GET", "/node")
        def node_list(ident, body, q):
            ident = app.require(ident)
            rows = db.query("SELECT * FROM node ORDER BY id")
            if ident["type"] == "user":
                reach = app._orgs_in_reach(ident, app.scope_of(ident, "node", "view"))
                rows = [n for n in rows if reach is None or n["organization_id"] in reach]
            else:
                rows = [n for n in rows if n["collaboration_id"] == ident["collaboration_id"]]
            return [app.node_json(n) for n in rows]

        @app.route("POST", "/node")
        def node_create(ident, body, q):
            ident = app.require(ident, "user")
            sc = app.scope_of(ident, "node", "create")
            if sc is None:
                raise HTTPError(401, "You lack the permission to do that!")
            cid = body.get("collaboration_id")
            c = db.get("collaboration", int(cid)) if cid else None
            if c is None:
                raise HTTPError(404, f"collaboration id={cid} does not exist")
            oid = int(body.get("organization_id") or ident["organization_id"])
            if sc != "global" and oid != ident["organization_id"]:
                raise HTTPError(401, "You are not allowed to create a node for another organization")
            if oid not in db.collaboration_organizations(c["id"]):
                raise HTTPError(400, f"organization id={oid} is not part of collaboration id={c['id']}")
            if db.one("SELECT id FROM node WHERE organization_id=? AND collaboration_id=?", (oid, c["id"])):
                raise HTTPError(400, "Your organization already has a node for this collaboration")
            org = db.get("organization", oid)
            name = body.get("name") or f"{org['name']} - {c['name']} Node"
            nid = db.insert("node", name=name, api_key=db.new_api_key(), collaboration_id=c["id"], organization_id=oid)
            return app.node_json(db.get("node", nid), with_key=True), 201

        @app.route("GET", r"/node/(\d+)")
        def node_get(ident, body, q, nid):
            ident = app.require(ident)
            n = db.get("node", int(nid))
            if n is None:
                raise HTTPError(404, f"node id={nid} is not found")
            if not (ident["type"] == "node" and ident["id"] == n["id"]):
                app.require_collaboration_view(ident, n["collaboration_id"], "node")
            return app.node_json(n)

        @app.route("PATCH", r"/node/(\d+)")
        def node_patch(ident, body, q, nid):
            ident = app.require(ident, "user", "node")
            n = db.get("node", int(nid))
            if n is None:
                raise HTTPError(404, f"node id={nid} is not found")
            if ident["type"] == "node" and ident["id"] != n["id"]:
                raise HTTPError(401, "A node can only update itself")
            if ident["type"] == "user":
                sc = app.scope_of(ident, "node", "edit")
                if not (sc == "global" or (sc and n["organization_id"] == ident["organization_id"])):
                    raise HTTPError(401, "You lack the permission to do that!")
            fields = {k: body[k] for k in ("name", "ip", "status", "gpu") if k in body}
            if fields.get("status") or ident["type"] == "node":
                fields["last_seen"] = now()
            db.update("node", n["id"], **fields)
            if "status" in fields:
                app.events.emit("node-status-changed", {"id": n["id"], "name": n["name"], "online": fields["status"] == "online"},
                                [f"collaboration_{n['collaboration_id']}"])
            return app.node_json(db.get("node", n["id"]))

        @app.route("DELETE", r"/node/(\d+)")
        def node_delete(ident, body, q, nid):
            ident = app.require(ident, "user")
            n = db.get("node", int(nid))
            if n is None:
                raise HTTPError(404, f"node id={nid} is not found")
            sc = app.scope_of(ident, "node", "delete")
            if not (sc == "global" or (sc and n["organization_id"] == ident["organization_id"])):
                raise HTTPError(401, "You lack the permission to do that!")
            db.delete("node", n["id"])
            return {"msg": f"Successfully deleted node id={nid}"}

        # ---- user / role / rule
        @app.route("GET", "/user")
        def user_list(ident, body, q):
            ident = app.require(ident, "user")
            reach = app._orgs_in_reach(ident, app.scope_of(ident, "user", "view"))
            rows = db.query("SELECT * FROM user ORDER BY id")
            return [app.user_json(u) for u in rows if reach is None or u["organization_id"] in reach or u["id"] == ident["id"]]

        @app.route("POST", "/user")
        def user_create(ident, body, q):
            ident = app.require(ident, "user")
            sc = app.scope_of(ident, "user", "create")
            if sc is None:
                raise HTTPError(401, "You lack the permission to do that!")
            for k in ("username", "password"):
                if not body.get(k):
                    raise HTTPError(400, f"{k} is required")
            if db.one("SELECT id FROM user WHERE username=?", (body["username"],)):
                raise HTTPError(400, "username already exists.")
            oid = int(body.get("organization_id") or ident["organization_id"])
            if sc != "global" and oid != ident["organization_id"]:
                raise HTTPError(401, "You lack the permission to create users for another organization")
            app.check_grant(ident, body.get("roles", []), body.get("rules", []), sc)
            uid = db.insert("user", username=body["username"], password=hash_password(body["password"]),
                            firstname=body.get("firstname"), lastname=body.get("lastname"), email=body.get("email"),
                            organization_id=oid)
            for rid in body.get("roles", []):
                db.execute("INSERT OR IGNORE INTO user_role VALUES (?,?)", (uid, int(rid)))
            for rid in body.get("rules", []):
                db.execute("INSERT OR IGNORE INTO user_rule VALUES (?,?)", (uid, int(rid)))
            return app.user_json(db.get("user", uid)), 201

        @app.route("GET", r"/user/(\d+)")
        def user_get(ident, body, q, uid):
            ident = app.require(ident, "user")
            u = db.get("user", int(uid))
            if u is None:
                raise HTTPError(404, f"user id={uid} is not found")
            reach = app._orgs_in_reach(ident, app.scope_of(ident, "user", "view"))
            if not (u["id"] == ident["id"] or reach is None or u["organization_id"] in reach):
                raise HTTPError(401, "You lack the permission to do that!")
            return app.user_json(u)

        @app.route("PATCH", r"/user/(\d+)")
        def user_patch(ident, body, q, uid):
            ident = app.require(ident, "user")
            u = db.get("user", int(uid))
            if u is None:
                raise HTTPError(404, f"user id={uid} not found")
            sc = app.scope_of(ident, "user", "edit")
            if not (u["id"] == ident["id"] or sc == "global" or (sc and u["organization_id"] == ident["organization_id"])):
                raise HTTPError(401, "You lack the permission to do that!")
            fields = {k: body[k] for k in ("firstname", "lastname", "email") if k in body}
            if body.get("password"):
                fields["password"] = hash_password(body["password"])
            if ("roles" in body or "rules" in body) and not sc:
                raise HTTPError(401, "You lack the permission to change roles or rules")
            if "roles" in body or "rules" in body:
                app.check_grant(ident, body.get("roles", []), body.get("rules", []), sc)     # before anything is written
            db.update("user", u["id"], **fields)
            if "roles" in body:
                db.execute("DELETE FROM user_role WHERE user_id=?", (u["id"],))
                for rid in body["roles"]:
                    db.execute("INSERT OR IGNORE INTO user_role VALUES (?,?)", (u["id"], int(rid)))
            if "rules" in body:
                db.execute("DELETE FROM user_rule WHERE user_id=?", (u["id"],))
                for rid in body["rules"]:
                    db.execute("INSERT OR IGNORE INTO user_rule VALUES (?,?)", (u["id"], int(rid)))
            return app.user_json(db.get("user", u["id"]))

        @app.route("DELETE", r"/user/(\d+)")
        def user_delete(ident, body, q, uid):
            ident = app.require(ident, "user")
            u = db.get("user", int(uid))
            if u is None:
                raise HTTPError(404, f"user id={uid} not found")
            sc = app.scope_of(ident, "user", "delete")
            if not (sc == "global" or (sc and u["organization_id"] == ident["organization_id"])):
                raise HTTPError(401, "You lack the permission to do that!")
            db.delete("user", u["id"])
            return {"msg": f"user id={uid} is removed from the database"}

        @app.route("GET", "/role")
        def role_list(ident, body, q):
            """Default roles plus the roles of the caller's organization (every role with global ``role view`` scope)."""
            ident = app.require(ident, "user")
            everything = app.scope_of(ident, "role", "view") == "global"
            return [admin_routes.role_json(app, r) for r in db.query("SELECT * FROM role ORDER BY id")
                    if everything or r["organization_id"] in (None, ident["organization_id"])]

        @app.route("GET", "/rule")
        def rule_list(ident, body, q):
            app.require(ident, "user")
            return db.query("SELECT * FROM rule ORDER BY id")

        # ---- task
        @app.route("GET", "/task")
        def task_list(ident, body, q):
            ident = app.require(ident)
            rows = db.query("SELECT * FROM task ORDER BY id")
            if ident["type"] == "user":
                sc = app.scope_of(ident, "task", "view")
                if sc != "global":
                    mine = set(db.organization_collaborations(ident["organization_id"])) if sc else set()
                    rows = [t for t in rows if t["collaboration_id"] in mine]
            else:
                rows = [t for t in rows if t["collaboration_id"] == ident["collaboration_id"]]
            for key in ("run_id", "parent_id", "collaboration_id"):
                if key in q:
                    rows = [t for t in rows if str(t[key]) == q[key]]
            if "image" in q:
                rows = [t for t in rows if t["image"] == q["image"]]
            return [app.task_json(t, include_results=q.get("include") == "results") for t in rows]

        @app.route("POST", "/task")
        def task_create(ident, body, q):
            ident = app.require(ident, "user", "container")
            cid = body.get("collaboration_id")
            c = db.get("collaboration", int(cid)) if cid is not None else None
            if c is None:
                raise HTTPError(404, f"Collaboration id={cid} not found!")
            orgs = body.get("organizations") or []
            if not orgs:
                raise HTTPError(400, "No organizations (with their input) specified")
            members = set(db.collaboration_organizations(c["id"]))
            for o in orgs:
                if int(o.get("id", -1)) not in members:
                    raise HTTPError(400, f"organization id={o.get('id')} is not part of collaboration id={c['id']}")
            image = body.get("image")
            if not image:
                raise HTTPError(400, "image is required")
            parent_id, run_id, initiator, init_user, parent_db = None, None, None, None, None
            if ident["type"] == "user":
                sc = app.scope_of(ident, "task", "create")
                if not (sc == "global" or (sc and ident["organization_id"] in members)):
                    raise HTTPError(401, "You lack the permission to do that!")
                initiator, init_user = ident["organization_id"], ident["id"]
                run_id = db.next_run_id()
            else:  # container: sub-task of its own task, same image, same collaboration
                if ident["collaboration_id"] != c["id"]:
                    raise HTTPError(401, "Container does not belong to the collaboration it is posting a task to")
                if ident["image"] != image:
                    raise HTTPError(401, f"Container does not have permission to use image {image!r}")
                parent = db.get("task", ident["task_id"])
                if parent is None or db.task_complete(parent["id"]):
                    raise HTTPError(401, "Parent task is finished: no new sub-tasks allowed")
                parent_id, run_id, initiator, init_user = parent["id"], parent["run_id"], ident["organization_id"], parent["init_user_id"]
                parent_db = parent["database"]             # a sub-task reads the database label its parent was pointed at
            tid = db.insert("task", name=body.get("name", ""), description=body.get("description", ""), image=image,
                            collaboration_id=c["id"], run_id=run_id, parent_id=parent_id,
                            database=body.get("database") or parent_db or "default", initiator_id=initiator, init_user_id=init_user,
                            created_at=now())
            for o in orgs:
                inp = o.get("input")
                rid = db.insert("result", task_id=tid, organization_id=int(o["id"]),
                                input=inp if isinstance(inp, str) else json.dumps(inp), assigned_at=now())
                node = db.one("SELECT id FROM node WHERE organization_id=? AND collaboration_id=?", (int(o["id"]), c["id"]))
                if node is None:
                    log.warning("organization %s has no node in collaboration %s", o["id"], c["id"])
                if node is not None and len(inp if isinstance(inp, str) else "") <= 65536:
                    # the node that has to run it gets the whole work item in its own room (nobody else listens there):
                    # no GET /result round trip before it can start
                    app.events.emit("new_task", {"task_id": tid, "result_id": rid, "organization_id": int(o["id"]),
                                                 "result": app.result_json(db.get("result", rid), with_task=True)},
                                    [f"node_{node['id']}"])
                app.events.emit("new_task", {"task_id": tid, "result_id": rid, "organization_id": int(o["id"])},
                                [f"collaboration_{c['id']}"])
            return app.task_json(db.get("task", tid)), 201

        @app.route("GET", r"/task/(\d+)")
        def task_get(ident, body, q, tid):
            ident = app.require(ident)
            t = db.get("task", int(tid))
            if t is None:
                raise HTTPError(404, f"task id={tid} is not found")
            app.require_collaboration_view(ident, t["collaboration_id"], "task")
            return app.task_json(t, include_results=q.get("include") == "results")

        @app.route("GET", r"/task/(\d+)/result")
        def task_results(ident, body, q, tid):
            ident = app.require(ident)
            t = db.get("task", int(tid))
            if t is None:
                raise HTTPError(404, f"task id={tid} is not found")
            app.require_collaboration_view(ident, t["collaboration_id"], "result")
            return [app.result_json(r) for r in db.query("SELECT * FROM result WHERE task_id=? ORDER BY id", (int(tid),))]

        @app.route("DELETE", r"/task/(\d+)")
        def task_delete(ident, body, q, tid):
            ident = app.require(ident, "user")
            t = db.get("task", int(tid))
            if t is None:
                raise HTTPError(404, f"task id={tid} not found")
            sc = app.scope_of(ident, "task", "delete")
            if not (sc == "global" or (sc and t["initiator_id"] == ident["organization_id"])):
                raise HTTPError(401, "You lack the permission to do that!")

            def rm(task_id):
                for ch in db.query("SELECT id FROM task WHERE parent_id=?", (task_id,)):
                    rm(ch["id"])
                db.execute("DELETE FROM result WHERE task_id=?", (task_id,))
                db.delete("task", task_id)
            rm(t["id"])
            app.events.emit("kill_containers", {"task_id": t["id"]}, [f"collaboration_{t['collaboration_id']}"])
            return {"msg": f"task id={tid} and its results successfully deleted"}

        # ---- result
        @app.route("GET", "/result")
        def result_list(ident, body, q):
            ident = app.require(ident)
            sql, args = "SELECT result.* FROM result JOIN task ON task.id = result.task_id WHERE 1=1", []
            if "task_id" in q:
                sql += " AND result.task_id=?"
                args.append(int(q["task_id"]))
            if "organization_id" in q:
                sql += " AND result.organization_id=?"
                args.append(int(q["organization_id"]))
            if "node_id" in q:
                n = db.get("node", int(q["node_id"]))
                if n is None:
                    raise HTTPError(404, f"node id={q['node_id']} not found")
                sql += " AND result.organization_id=? AND task.collaboration_id=?"
                args += [n["organization_id"], n["collaboration_id"]]
            if q.get("state") == "open":
                sql += " AND result.finished_at IS NULL"
            if ident["type"] in ("node", "container"):
                sql += " AND task.collaboration_id=?"
                args.append(ident["collaboration_id"])
            elif app.scope_of(ident, "result", "view") != "global":
                mine = db.organization_collaborations(ident["organization_id"]) if app.scope_of(ident, "result", "view") else []
                sql += f" AND task.collaboration_id IN ({','.join('?' * len(mine)) or 'NULL'})"
                args += mine
            rows = db.query(sql + " ORDER BY result.id", args)
            return [app.result_json(r, with_task=q.get("include") == "task") for r in rows]

        @app.route("GET", r"/result/(\d+)")
        def result_get(ident, body, q, rid):
            ident = app.require(ident)
            r = db.get("result", int(rid))
            if r is None:
                raise HTTPError(404, f"result id={rid} not found")
            t = db.get("task", r["task_id"])
            app.require_collaboration_view(ident, t["collaboration_id"] if t else -1, "result")
            return app.result_json(r, with_task=q.get("include") == "task")

        @app.route("PATCH", r"/result/(\d+)")
        def result_patch(ident, body, q, rid):
            ident = app.require(ident, "node")
            r = db.get("result", int(rid))
            if r is None:
                raise HTTPError(404, f"result id={rid} not found")
            t = db.get("task", r["task_id"])
            if r["organization_id"] != ident["organization_id"] or t["collaboration_id"] != ident["collaboration_id"]:
                raise HTTPError(401, "This result does not belong to your organization/collaboration")
            if r["finished_at"] is not None:
                raise HTTPError(400, "Cannot update an already finished result!")
            fields = {k: body[k] for k in ("started_at", "finished_at", "result", "log", "status") if k in body}
            db.update("result", r["id"], **fields)
            if "finished_at" in fields or "status" in fields:
                app.events.emit("status_update", {"result_id": r["id"], "task_id": t["id"], "status": fields.get("status", "completed"),
                                                  "organization_id": r["organization_id"], "parent_id": t["parent_id"],
                                                  "task_complete": "finished_at" in fields and db.task_complete(t["id"])},
                                [f"collaboration_{t['collaboration_id']}", f"task_{t['id']}"])
            return app.result_json(db.get("result", r["id"]))

        # ---- events (long poll)
        @app.route("GET", "/event")
        def event_poll(ident, body, q):
            ident = app.require(ident)
            since = int(q.get("since", app.events.last_id()))
            timeout = min(float(q.get("timeout", 25)), 55.0)
            rooms = app.event_rooms(ident, q.get("task_id"))
            evs = app.events.wait(since, rooms, timeout)
            return {"events": evs, "last_id": evs[-1]["id"] if evs else since}

        admin_routes.register(app)          # roles / rules, account recovery, membership sub-resources

    # ------------------------------------------------------------------ http plumbing
    def make_handler(self):
        app = self

        class Handler(BaseHTTPRequestHandler):
            protocol_version = "HTTP/1.1"
            wbufsize = 64 * 1024                 # one send per response (no Nagle / delayed-ACK stall)
            disable_nagle_algorithm = True
            server_version = "vantage6-b200"

            def log_message(self, fmt, *args):
                log.debug("%s - %s", self.address_string(), fmt % args)

            def setup(self):
                super().setup()
                with app._conns_lock:
                    app._conns.add(self.connection)

            def finish(self):
                with app._conns_lock:
                    app._conns.discard(self.connection)
                super().finish()

            def _serve(self, method):
                parts = urlsplit(self.path)
                try:
                    length = int(self.headers.get("Content-Length") or 0)
                except ValueError:
                    length = -1
                if length < 0 or length > MAX_BODY_BYTES:      # refuse before reading: the framing of this connection is void
                    self.close_connection = True
                    data = json.dumps({"msg": f"request body must be 0..{MAX_BODY_BYTES} bytes with a valid Content-Length"}).encode("utf-8")
                    self.send_response(413 if length > 0 else 400)
                    self.send_header("Content-Type", "application/json")
                    self.send_header("Content-Length", str(len(data)))
                    self.send_header("Connection", "close")
                    self.end_headers()
                    self.wfile.write(data)
                    return
                raw = self.rfile.read(length) if length else b""
                try:
                    body = json.loads(raw.decode("utf-8")) if raw else {}
                except Exception:  # noqa: BLE001
                    body = {}
                t0 = time.perf_counter()
                status, payload = app.dispatch(method, parts.path, parse_qs(parts.query), body, self.headers)
                data = payload.encode("utf-8") if isinstance(payload, PlainText) else json.dumps(payload).encode("utf-8")
                if TRACE_HTTP:
                    log.info("http %s %s -> %s in %.2f ms", method, self.path[:80], status, 1e3 * (time.perf_counter() - t0))
                self.send_response(status)
                self.send_header("Content-Type", "text/plain; version=0.0.4; charset=utf-8" if isinstance(payload, PlainText) else "application/json")
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

            def do_PUT(self):
                self._serve("PUT")

        return Handler

    def reap_silent_nodes(self, timeout_s: Optional[float] = None) -> int:
        """Failure detection: a node whose last heartbeat is older than
        ``node_timeout_s`` is marked offline and the collaboration is told.  Returns #nodes reaped."""
        timeout_s = float(self.config.get("node_timeout_s", 60)) if timeout_s is None else timeout_s
        cutoff = (_dt.datetime.now(_dt.timezone.utc) - _dt.timedelta(seconds=timeout_s)).isoformat()
        n = 0
        for node in self.db.query("SELECT * FROM node WHERE status='online' AND (last_seen IS NULL OR last_seen < ?)", (cutoff,)):
            self.db.update("node", node["id"], status="offline")
            self.events.emit("node-status-changed", {"id": node["id"], "name": node["name"], "online": False},
                             [f"collaboration_{node['collaboration_id']}"])
            log.warning("node %s (%s) missed its heartbeats: marked offline", node["id"], node["name"])
            n += 1
        return n

    def _reaper_loop(self) -> None:
        while self._httpd is not None:
            time.sleep(5.0)
            try:
                self.reap_silent_nodes()
            except Exception:  # noqa: BLE001
                log.debug("reaper error", exc_info=True)

    def tls_context(self):
        """Server-side TLS from the ``ssl: {certfile, keyfile}`` block of the configuration (REST and the websocket event
        channel alike); ``None`` = plain HTTP, e.g. behind a reverse proxy that terminates TLS, as vantage6 deployments do."""
        cfg = self.config.get("ssl") or {}
        if not cfg.get("certfile"):
            return None
        import ssl

        ctx = ssl.SSLContext(ssl.PROTOCOL_TLS_SERVER)
        ctx.minimum_version = ssl.TLSVersion.TLSv1_2
        ctx.load_cert_chain(cfg["certfile"], cfg.get("keyfile"))
        return ctx

    def start(self, ip: str = "127.0.0.1", port: int = 5000, block: bool = False) -> int:
        """Serve; returns the bound port (``port=0`` picks a free one)."""
        threading.Thread(target=lambda: (time.sleep(1.0), self._reaper_loop()), daemon=True).start()
        self._httpd = _ApiHTTPServer((ip, port), self.make_handler())
        bound = self._httpd.server_address[1]
        tls = self.tls_context()
        if tls is not None:     # handshake in the connection's own thread (first read), not in the accept loop
            self._httpd.socket = tls.wrap_socket(self._httpd.socket, server_side=True, do_handshake_on_connect=False)
        if self.config.get("event_websocket", True):
            try:
                from .ws_events import WebSocketEvents

                ws_port = int(self.config.get("event_port") or 0)
                self.ws = WebSocketEvents(self, ip, ws_port, ssl_context=tls)
                self.ws.start()
                self.events.push = self.ws.publish
            except Exception as e:  # noqa: BLE001 -- no `websockets` package: long-poll only
                log.warning("websocket event channel unavailable (%s): long-poll only", e)
                self.ws = None
        log.info("vantage6-b200 server %s listening on %s://%s:%s%s", __version__, "https" if tls else "http", ip, bound, self.api_path)
        if block:
            try:
                self._httpd.serve_forever(poll_interval=0.2)
            finally:
                self._httpd.server_close()
        else:
            self._thread = threading.Thread(target=self._httpd.serve_forever, kwargs={"poll_interval": 0.2}, daemon=True)
            self._thread.start()
        return bound

    def stop(self) -> None:
        if getattr(self, "ws", None) is not None:
            self.events.push = None
            self.ws.stop()
            self.ws = None
        if self._httpd is not None:
            self._httpd.shutdown()
            self._httpd.server_close()
            self._httpd = None
            with self._conns_lock:                 # kept-alive connections would otherwise be served on by their threads
                conns = list(self._conns)
            for conn in conns:
                try:
                    conn.shutdown(socket.SHUT_RDWR)
                except OSError:
                    pass
