"""Second half of the server's route table: role / rule management, account recovery and the membership
sub-resources (the parts of the vantage6 3.x REST surface that administer the entities rather than run tasks;
SURVEY.md Appendix C).  ``register(app)`` is called from ``ServerApp._register_routes``.

* ``/role`` -- custom roles per organization: ``POST``, ``GET/PATCH/DELETE /role/<id>``, ``GET /role/<id>/rule``,
  ``POST/DELETE /role/<id>/rule/<rule_id>``; ``GET /rule/<id>``.  The rule that protects ``/user`` protects these too:
  nobody can put a permission into a role that they do not hold themselves (``ServerApp.check_grant``), default roles
  (no organization) are read-only below global scope and cannot be deleted at all.
* ``/recover/lost``, ``/recover/reset``, ``/password/change`` -- password recovery with a short-lived signed reset token
  (mailed when ``smtp`` is configured, otherwise parked in the running server's ``app.outbox``; the operator can also
  mint one from ``vserver shell`` with ``reset_token(username)``).
* ``/collaboration/<id>/organization`` and ``/collaboration/<id>/node`` (``POST`` / ``DELETE``), ``/organization/<id>/
  collaboration``, ``/organization/<id>/node``, ``/node/<id>/task``.
* ``/port`` -- the address book of running algorithms (vantage6: the VPN ports of algorithm containers): a node registers
  under which address / port / label the algorithm it runs for a result can be reached by its siblings of the same run.
  On one NVSwitch box the "address" is the node's GPU and its rank in the run's rendezvous (algorithm/peer.py).
* ``/spec`` -- the route table itself.  ``/metrics`` -- counters and gauges in the Prometheus text format.
"""
from __future__ import annotations

import datetime as _dt
import json
import logging
import smtplib
from email.message import EmailMessage

import jwt

from .db import check_password, hash_password, now

log = logging.getLogger("server")

RESET_TOKEN_MINUTES = 60
LOST_REPLY = {"msg": "If the username or email is in our database you will soon receive an email"}


def role_json(app, r: dict) -> dict:
    db = app.db
    return {"id": r["id"], "name": r["name"], "description": r["description"],
            "organization": app.link("organization", r["organization_id"]) if r["organization_id"] else None,
            "organization_id": r["organization_id"],
            "rules": [app.link("rule", x["rule_id"]) for x in db.query("SELECT rule_id FROM role_rule WHERE role_id=? ORDER BY rule_id", (r["id"],))],
            "users": [app.link("user", x["user_id"]) for x in db.query("SELECT user_id FROM user_role WHERE role_id=? ORDER BY user_id", (r["id"],))]}


def deliver_reset_token(app, user: dict, token: str) -> None:
    """Mail the token when the server has an ``smtp`` block (``server``, ``port``, optional ``username`` /
    ``password`` / ``email_from``); otherwise keep it in ``app.outbox`` (never in the log) for the operator."""
    smtp = app.config.get("smtp") or {}
    if smtp.get("server") and user.get("email"):
        msg = EmailMessage()
        msg["Subject"] = "password reset"
        msg["From"] = smtp.get("email_from", "noreply@vantage6.local")
        msg["To"] = user["email"]
        msg.set_content(f"Dear {user.get('firstname') or user['username']},\n\nyour password reset token (valid for "
                        f"{RESET_TOKEN_MINUTES} minutes):\n\n{token}\n")
        try:
            with smtplib.SMTP(smtp["server"], int(smtp.get("port", 25)), timeout=10) as s:
                if smtp.get("username"):
                    s.starttls()
                    s.login(smtp["username"], smtp.get("password", ""))
                s.send_message(msg)
            return
        except Exception as e:  # noqa: BLE001
            log.warning("could not mail the reset token of user id=%s (%s); parked in the outbox", user["id"], e)
    else:
        log.warning("no smtp server configured: reset token of user id=%s parked in the outbox", user["id"])
    app.outbox.append({"user_id": user["id"], "username": user["username"], "email": user.get("email"),
                       "reset_token": token, "created_at": now()})
    del app.outbox[:-100]


def make_reset_token(app, user: dict) -> str:
    exp = _dt.datetime.now(_dt.timezone.utc) + _dt.timedelta(minutes=RESET_TOKEN_MINUTES)
    # the tail of the current password hash is part of the claim: a token dies with the password it was issued for
    payload = {"sub": json.dumps({"id": user["id"], "pw": user["password"][-16:]}), "exp": exp, "typ": "reset"}
    return jwt.encode(payload, app.secret, algorithm="HS256")


def issue_reset_token(app, username: str) -> str:
    """For the operator (``vserver shell``: ``reset_token("alice")``) on a server without mail."""
    u = app.db.one("SELECT * FROM user WHERE username=?", (username,))
    if u is None:
        raise KeyError(f"no user {username!r}")
    return make_reset_token(app, u)


def register(app) -> None:  # noqa: C901 -- a flat route table reads best in one place
    from .app import HTTPError          # late: app.py imports this module

    db = app.db
    app.outbox = []

    # ------------------------------------------------------------------ roles
    def visible_role(ident, r) -> bool:
        sc = app.scope_of(ident, "role", "view")
        return r["organization_id"] is None or sc == "global" or r["organization_id"] == ident["organization_id"]

    def editable_role(ident, r, operation: str) -> None:
        sc = app.scope_of(ident, "role", operation)
        if sc is None:
            raise HTTPError(401, "You lack the permission to do that!")
        if r["organization_id"] is None:
            if sc != "global":
                raise HTTPError(401, "Default roles can only be changed with global scope")
        elif sc != "global" and r["organization_id"] != ident["organization_id"]:
            raise HTTPError(401, "You cannot change a role of another organization")

    def get_role(rid) -> dict:
        r = db.get("role", int(rid))
        if r is None:
            raise HTTPError(404, f"role id={rid} not found")
        return r

    @app.route("GET", r"/role/(\d+)")
    def role_get(ident, body, q, rid):
        ident = app.require(ident, "user")
        r = get_role(rid)
        if not visible_role(ident, r):
            raise HTTPError(401, "You do not have permission to view this role")
        return role_json(app, r)

    @app.route("POST", "/role")
    def role_create(ident, body, q):
        ident = app.require(ident, "user")
        sc = app.scope_of(ident, "role", "create")
        if sc is None:
            raise HTTPError(401, "You lack the permission to do that!")
        if not body.get("name"):
            raise HTTPError(400, "name is required")
        oid = body.get("organization_id", ident["organization_id"])
        if oid is not None:
            oid = int(oid)
            if db.get("organization", oid) is None:
                raise HTTPError(404, f"organization id={oid} not found")
        if sc != "global" and oid != ident["organization_id"]:
            raise HTTPError(401, "You cannot create a role for another organization")
        rules = [int(x) for x in body.get("rules", [])]
        app.check_grant(ident, [], rules, sc)
        if db.one("SELECT id FROM role WHERE name=? AND organization_id IS ?", (body["name"], oid)):
            raise HTTPError(400, f"role {body['name']!r} already exists in this organization")
        rid = db.insert("role", name=body["name"], description=body.get("description"), organization_id=oid)
        for x in rules:
            db.execute("INSERT OR IGNORE INTO role_rule VALUES (?,?)", (rid, x))
        return role_json(app, db.get("role", rid)), 201

    @app.route("PATCH", r"/role/(\d+)")
    def role_patch(ident, body, q, rid):
        ident = app.require(ident, "user")
        r = get_role(rid)
        editable_role(ident, r, "edit")
        if "rules" in body:
            rules = [int(x) for x in body["rules"]]
            app.check_grant(ident, [], rules, app.scope_of(ident, "role", "edit"))      # before anything is written
        db.update("role", r["id"], **{k: body[k] for k in ("name", "description") if k in body})
        if "rules" in body:
            db.execute("DELETE FROM role_rule WHERE role_id=?", (r["id"],))
            for x in rules:
                db.execute("INSERT OR IGNORE INTO role_rule VALUES (?,?)", (r["id"], x))
        return role_json(app, db.get("role", r["id"]))

    @app.route("DELETE", r"/role/(\d+)")
    def role_delete(ident, body, q, rid):
        ident = app.require(ident, "user")
        r = get_role(rid)
        if r["organization_id"] is None:
            raise HTTPError(400, "Default roles cannot be deleted")
        editable_role(ident, r, "delete")
        holders = db.query("SELECT user_id FROM user_role WHERE role_id=?", (r["id"],))
        if holders and str(q.get("delete_dependents", body.get("delete_dependents", ""))).lower() not in ("1", "true", "yes"):
            raise HTTPError(400, f"role id={rid} is assigned to {len(holders)} user(s); pass delete_dependents=true to revoke it from them")
        db.execute("DELETE FROM user_role WHERE role_id=?", (r["id"],))
        db.execute("DELETE FROM role_rule WHERE role_id=?", (r["id"],))
        db.delete("role", r["id"])
        return {"msg": f"role id={rid} successfully deleted"}

    @app.route("GET", r"/role/(\d+)/rule")
    def role_rules(ident, body, q, rid):
        ident = app.require(ident, "user")
        r = get_role(rid)
        if not visible_role(ident, r):
            raise HTTPError(401, "You do not have permission to view this role")
        return db.query("SELECT rule.* FROM rule JOIN role_rule ON rule.id = role_rule.rule_id WHERE role_rule.role_id=? ORDER BY rule.id", (r["id"],))

    @app.route("POST", r"/role/(\d+)/rule/(\d+)")
    def role_rule_add(ident, body, q, rid, rule_id):
        ident = app.require(ident, "user")
        r = get_role(rid)
        editable_role(ident, r, "edit")
        app.check_grant(ident, [], [int(rule_id)], app.scope_of(ident, "role", "edit"))
        db.execute("INSERT OR IGNORE INTO role_rule VALUES (?,?)", (r["id"], int(rule_id)))
        return role_json(app, r), 201

    @app.route("DELETE", r"/role/(\d+)/rule/(\d+)")
    def role_rule_remove(ident, body, q, rid, rule_id):
        ident = app.require(ident, "user")
        r = get_role(rid)
        editable_role(ident, r, "edit")
        if db.one("SELECT 1 AS x FROM role_rule WHERE role_id=? AND rule_id=?", (r["id"], int(rule_id))) is None:
            raise HTTPError(404, f"rule id={rule_id} is not part of role id={rid}")
        db.execute("DELETE FROM role_rule WHERE role_id=? AND rule_id=?", (r["id"], int(rule_id)))
        return role_json(app, r)

    @app.route("GET", r"/rule/(\d+)")
    def rule_get(ident, body, q, rule_id):
        app.require(ident, "user")
        rule = db.get("rule", int(rule_id))
        if rule is None:
            raise HTTPError(404, f"rule id={rule_id} not found")
        return rule

    # ------------------------------------------------------------------ account recovery
    def check_new_password(pw) -> str:
        if not isinstance(pw, str) or len(pw) < int((app.config.get("password_policy") or {}).get("min_length", 4)):
            raise HTTPError(400, "The new password is too short")
        return pw

    @app.route("POST", "/recover/lost")
    def recover_lost(ident, body, q):
        username, email = body.get("username"), body.get("email")
        if not username and not email:
            raise HTTPError(400, "No username or email provided!")
        u = db.one("SELECT * FROM user WHERE username=?", (username,)) if username else db.one("SELECT * FROM user WHERE email=?", (email,))
        if u is not None:           # the reply is the same either way: no account enumeration
            deliver_reset_token(app, u, make_reset_token(app, u))
        return LOST_REPLY

    @app.route("POST", "/recover/reset")
    def recover_reset(ident, body, q):
        token, password = body.get("reset_token"), body.get("password")
        if not token or not password:
            raise HTTPError(400, "The reset token and/or password is missing!")
        try:
            payload = jwt.decode(token, app.secret, algorithms=["HS256"])
            claim = json.loads(payload["sub"])
            if payload.get("typ") != "reset":
                raise jwt.InvalidTokenError("not a reset token")
        except jwt.PyJWTError:
            raise HTTPError(401, "Invalid or expired recovery token!")
        u = db.get("user", int(claim["id"]))
        if u is None or u["password"][-16:] != claim.get("pw"):
            raise HTTPError(401, "Invalid or expired recovery token!")
        db.update("user", u["id"], password=hash_password(check_new_password(password)), failed_login_attempts=0)
        return {"msg": "The password has successfully been reset!"}

    @app.route("PATCH", "/password/change")
    def password_change(ident, body, q):
        ident = app.require(ident, "user")
        cur, new = body.get("current_password"), body.get("new_password")
        if not cur or not new:
            raise HTTPError(400, "current_password and new_password are required")
        u = db.get("user", ident["id"])
        if u is None or not check_password(cur, u["password"]):
            raise HTTPError(401, "Your current password is not correct!")
        if cur == new:
            raise HTTPError(400, "New password is the same as current password!")
        db.update("user", u["id"], password=hash_password(check_new_password(new)))
        return {"msg": "The password has been changed successfully!"}

    # ------------------------------------------------------------------ membership sub-resources
    def get_collab(cid) -> dict:
        c = db.get("collaboration", int(cid))
        if c is None:
            raise HTTPError(404, f"collaboration id={cid} can not be found")
        return c

    def require_collab_edit(ident, c) -> None:
        sc = app.scope_of(ident, "collaboration", "edit")
        if not (sc == "global" or (sc == "collaboration" and c["id"] in db.organization_collaborations(ident["organization_id"]))):
            raise HTTPError(401, "You lack the permission to do that!")

    @app.route("POST", r"/collaboration/(\d+)/organization")
    def collab_org_add(ident, body, q, cid):
        ident = app.require(ident, "user")
        c = get_collab(cid)
        require_collab_edit(ident, c)
        o = db.get("organization", int(body["id"])) if body.get("id") is not None else None
        if o is None:
            raise HTTPError(404, f"organization with id={body.get('id')} not found")
        db.execute("INSERT OR IGNORE INTO member VALUES (?,?)", (c["id"], o["id"]))
        return [app.link("organization", x) for x in db.collaboration_organizations(c["id"])]

    @app.route("DELETE", r"/collaboration/(\d+)/organization")
    def collab_org_remove(ident, body, q, cid):
        ident = app.require(ident, "user")
        c = get_collab(cid)
        require_collab_edit(ident, c)
        oid = body.get("id", q.get("id"))
        if oid is None or int(oid) not in db.collaboration_organizations(c["id"]):
            raise HTTPError(404, f"organization with id={oid} is not part of collaboration id={cid}")
        if db.one("SELECT id FROM node WHERE organization_id=? AND collaboration_id=?", (int(oid), c["id"])):
            raise HTTPError(400, "The organization still has a node in this collaboration: delete or detach the node first")
        db.execute("DELETE FROM member WHERE collaboration_id=? AND organization_id=?", (c["id"], int(oid)))
        return [app.link("organization", x) for x in db.collaboration_organizations(c["id"])]

    @app.route("POST", r"/collaboration/(\d+)/node")
    def collab_node_add(ident, body, q, cid):
        """Attach an existing (detached, or to be moved) node to this collaboration."""
        ident = app.require(ident, "user")
        c = get_collab(cid)
        require_collab_edit(ident, c)
        n = db.get("node", int(body["id"])) if body.get("id") is not None else None
        if n is None:
            raise HTTPError(404, f"node id={body.get('id')} not found")
        if n["collaboration_id"] == c["id"]:
            raise HTTPError(400, f"node id={n['id']} is already in collaboration id={cid}")
        if n["collaboration_id"] is not None:          # moving a node takes it away from where it is: needs the say there too
            require_collab_edit(ident, get_collab(n["collaboration_id"]))
        if n["organization_id"] not in db.collaboration_organizations(c["id"]):
            raise HTTPError(400, f"the node's organization id={n['organization_id']} is not part of collaboration id={cid}")
        if db.one("SELECT id FROM node WHERE organization_id=? AND collaboration_id=?", (n["organization_id"], c["id"])):
            raise HTTPError(400, "The organization already has a node in this collaboration")
        db.update("node", n["id"], collaboration_id=c["id"], status="offline")
        return [app.link("node", x["id"]) for x in db.query("SELECT id FROM node WHERE collaboration_id=? ORDER BY id", (c["id"],))], 201

    @app.route("DELETE", r"/collaboration/(\d+)/node")
    def collab_node_remove(ident, body, q, cid):
        """Detach a node: it keeps its api key but cannot authenticate until it is attached again."""
        ident = app.require(ident, "user")
        c = get_collab(cid)
        require_collab_edit(ident, c)
        nid = body.get("id", q.get("id"))
        n = db.get("node", int(nid)) if nid is not None else None
        if n is None or n["collaboration_id"] != c["id"]:
            raise HTTPError(404, f"node id={nid} is not part of collaboration id={cid}")
        db.update("node", n["id"], collaboration_id=None, status="offline")
        app.events.emit("node-status-changed", {"id": n["id"], "name": n["name"], "online": False}, [f"collaboration_{c['id']}"])
        return [app.link("node", x["id"]) for x in db.query("SELECT id FROM node WHERE collaboration_id=? ORDER BY id", (c["id"],))]

    def require_org_view(ident, oid: int) -> dict:
        o = db.get("organization", oid)
        if o is None:
            raise HTTPError(404, f"Organization id={oid} not found")
        if ident["type"] == "user":
            reach = app._orgs_in_reach(ident, app.scope_of(ident, "organization", "view"))
            if reach is not None and oid not in reach and oid != ident["organization_id"]:
                raise HTTPError(401, "You do not have permission to view this organization")
        elif oid not in db.collaboration_organizations(ident["collaboration_id"]):
            raise HTTPError(401, "You do not have permission to view this organization")
        return o

    @app.route("GET", r"/organization/(\d+)/collaboration")
    def org_collabs(ident, body, q, oid):
        ident = app.require(ident)
        o = require_org_view(ident, int(oid))
        cids = db.organization_collaborations(o["id"])
        return [app.collab_json(db.get("collaboration", c)) for c in cids if app.can_view_collaboration(ident, c, "collaboration")]

    @app.route("GET", r"/organization/(\d+)/node")
    def org_nodes(ident, body, q, oid):
        ident = app.require(ident)
        o = require_org_view(ident, int(oid))
        rows = db.query("SELECT * FROM node WHERE organization_id=? ORDER BY id", (o["id"],))
        return [app.node_json(n) for n in rows if n["collaboration_id"] is not None and app.can_view_collaboration(ident, n["collaboration_id"], "node")]

    @app.route("GET", r"/node/(\d+)/task")
    def node_tasks(ident, body, q, nid):
        """Tasks that have a result assigned to this node (``?state=open`` only those it still has to run)."""
        ident = app.require(ident)
        n = db.get("node", int(nid))
        if n is None:
            raise HTTPError(404, f"node id={nid} is not found")
        if not (ident["type"] == "node" and ident["id"] == n["id"]):
            app.require_collaboration_view(ident, n["collaboration_id"] if n["collaboration_id"] is not None else -1, "task")
        sql = ("SELECT DISTINCT task.* FROM task JOIN result ON result.task_id = task.id "
               "WHERE task.collaboration_id=? AND result.organization_id=?")
        if q.get("state") == "open":
            sql += " AND result.finished_at IS NULL"
        return [app.task_json(t) for t in db.query(sql + " ORDER BY task.id", (n["collaboration_id"], n["organization_id"]))]

    # ------------------------------------------------------------------ algorithm address book
    def port_json(p: dict) -> dict:
        r = db.get("result", p["result_id"])
        n = None
        if r is not None:
            t = db.get("task", r["task_id"])
            n = db.one("SELECT id, gpu, ip FROM node WHERE organization_id=? AND collaboration_id=?", (r["organization_id"], t["collaboration_id"])) if t else None
        return {"id": p["id"], "port": p["port"], "label": p["label"], "address": p["address"] or (n or {}).get("ip"),
                "result": app.link("result", p["result_id"]), "organization_id": r["organization_id"] if r else None,
                "node_id": n["id"] if n else None, "gpu": n["gpu"] if n else None}

    @app.route("POST", "/port")
    def port_create(ident, body, q):
        ident = app.require(ident, "node")
        for k in ("port", "result_id"):
            if body.get(k) is None:
                raise HTTPError(400, f"{k} is required")
        r = db.get("result", int(body["result_id"]))
        t = db.get("task", r["task_id"]) if r else None
        if r is None or t is None:
            raise HTTPError(404, f"result id={body['result_id']} not found")
        if r["organization_id"] != ident["organization_id"] or t["collaboration_id"] != ident["collaboration_id"]:
            raise HTTPError(401, "You lack the permissions to do that")
        pid = db.insert("port", result_id=r["id"], port=int(body["port"]), label=body.get("label"), address=body.get("address"))
        return port_json(db.get("port", pid)), 201

    @app.route("GET", "/port")
    def port_list(ident, body, q):
        """``?result_id=`` one algorithm, ``?task_id=`` every algorithm of a task, ``?run_id=`` of a whole run."""
        ident = app.require(ident)
        sql = ("SELECT port.*, task.collaboration_id AS cid FROM port JOIN result ON result.id = port.result_id "
               "JOIN task ON task.id = result.task_id WHERE 1=1")
        args = []
        for key, col in (("result_id", "port.result_id"), ("task_id", "result.task_id"), ("run_id", "task.run_id")):
            if key in q:
                sql += f" AND {col}=?"
                args.append(int(q[key]))
        rows = db.query(sql + " ORDER BY port.id", args)
        return [port_json(p) for p in rows if app.can_view_collaboration(ident, p["cid"], "port")]

    @app.route("DELETE", "/port")
    def port_delete(ident, body, q):
        ident = app.require(ident, "node")
        rid = q.get("result_id", body.get("result_id"))
        if rid is None:
            raise HTTPError(400, "result_id is required")
        r = db.get("result", int(rid))
        if r is None or r["organization_id"] != ident["organization_id"]:
            raise HTTPError(401, "You lack the permissions to do that")
        db.execute("DELETE FROM port WHERE result_id=?", (r["id"],))
        return {"msg": f"ports of result id={rid} removed"}

    # ------------------------------------------------------------------ the API describes itself
    @app.route("GET", "/spec")
    def spec(ident, body, q):
        out = []
        for method, rx, fn in app._routes:
            path = rx.pattern[1:-3].replace("(\\d+)", "<id>")
            out.append({"method": method, "path": app.api_path + path, "doc": " ".join((fn.__doc__ or "").split()) or None})
        return sorted(out, key=lambda r: (r["path"], r["method"]))

    # ------------------------------------------------------------------ observability
    @app.route("GET", "/metrics")
    def metrics(ident, body, q):
        """Prometheus text exposition: requests by route and status, handling time, events, listeners, nodes, tasks, results."""
        if not app.config.get("metrics_public", False):
            app.require(ident)
        from .app import PlainText

        def esc(v) -> str:
            return str(v).replace("\\", "\\\\").replace('"', '\\"')

        out = ["# HELP v6_http_requests_total Requests handled, by method, route and status.", "# TYPE v6_http_requests_total counter"]
        with app._stats_lock:
            stats = {k: list(v) for k, v in app._stats.items()}
        for (method, route, status), (n, _) in sorted(stats.items()):
            out.append(f'v6_http_requests_total{{method="{method}",route="{esc(route)}",status="{status}"}} {int(n)}')
        out += ["# HELP v6_http_request_seconds_total Time spent handling requests (long polls excluded).", "# TYPE v6_http_request_seconds_total counter"]
        per_route: dict = {}
        for (method, route, _), (n, sec) in stats.items():
            acc = per_route.setdefault((method, route), 0.0)
            per_route[(method, route)] = acc + sec
        for (method, route), sec in sorted(per_route.items()):
            out.append(f'v6_http_request_seconds_total{{method="{method}",route="{esc(route)}"}} {sec:.6f}')
        import time as _time

        gauges = [("v6_uptime_seconds", "Seconds since the server started.", _time.time() - app.started_at),
                  ("v6_events_emitted_total", "Events emitted on the event bus.", app.events.last_id()),
                  ("v6_event_listeners", "Open websocket event subscriptions.", app.ws.connections if app.ws is not None else 0),
                  ("v6_tasks_total", "Tasks in the database.", db.one("SELECT COUNT(*) AS n FROM task")["n"]),
                  ("v6_tasks_open", "Tasks with at least one unfinished result.",
                   db.one("SELECT COUNT(DISTINCT task_id) AS n FROM result WHERE finished_at IS NULL")["n"])]
        for name, help_, value in gauges:
            kind = "counter" if name.endswith("_total") and name != "v6_tasks_total" else "gauge"
            out += [f"# HELP {name} {help_}", f"# TYPE {name} {kind}", f"{name} {value}"]
        out += ["# HELP v6_nodes Nodes by status.", "# TYPE v6_nodes gauge"]
        for r in db.query("SELECT status, COUNT(*) AS n FROM node GROUP BY status ORDER BY status"):
            out.append(f'v6_nodes{{status="{esc(r["status"])}"}} {r["n"]}')
        out += ["# HELP v6_results Results by status.", "# TYPE v6_results gauge"]
        for r in db.query("SELECT status, COUNT(*) AS n FROM result GROUP BY status ORDER BY status"):
            out.append(f'v6_results{{status="{esc(r["status"])}"}} {r["n"]}')
        return PlainText("\n".join(out) + "\n")
