"""Bulk import of organizations / users / collaborations / nodes / tasks from a YAML file --
what ``vserver import`` runs inside the server runtime (``vserver-local import``: reference
vantage6/cli/server.py:487-489; the commented sketch ``fixture.load(entities, drop_all)`` at
reference server.py:518-523).  "Especially useful for testing" (reference server.py:418-421).

File format (vantage6 3.x)::

    organizations:
      - name: IKNL
        domain: iknl.nl
        public_key: <b64, optional>
        users:
          - {username: admin, firstname: a, lastname: b, password: secret, roles: [Root]}
    collaborations:
      - name: demo
        encrypted: false
        participants:
          - {name: IKNL, api-key: 123e4567-...}
        tasks:
          - {name: t1, image: v6-average-py, input: {...}}
"""
from __future__ import annotations

import json
import logging
from typing import Dict

from .db import Database, hash_password, now

log = logging.getLogger("fixtures")


def load(db: Database, entities: dict, drop_all: bool = False, ensure_defaults=None) -> Dict[str, int]:
    counts = {"organizations": 0, "users": 0, "collaborations": 0, "nodes": 0, "tasks": 0}
    if drop_all:
        db.drop_all()
        if ensure_defaults is not None:
            ensure_defaults()
    org_ids: Dict[str, int] = {}
    for org in entities.get("organizations", []) or []:
        existing = db.one("SELECT id FROM organization WHERE name=?", (org["name"],))
        if existing:
            oid = existing["id"]
        else:
            oid = db.insert("organization", name=org["name"], domain=org.get("domain"), address1=org.get("address1"),
                            address2=org.get("address2"), zipcode=str(org.get("zipcode", "")) or None,
                            country=org.get("country"), public_key=org.get("public_key"))
            counts["organizations"] += 1
        org_ids[org["name"]] = oid
        for usr in org.get("users", []) or []:
            if db.one("SELECT id FROM user WHERE username=?", (usr["username"],)):
                continue
            uid = db.insert("user", username=usr["username"], password=hash_password(str(usr.get("password", ""))),
                            firstname=usr.get("firstname"), lastname=usr.get("lastname"), email=usr.get("email"),
                            organization_id=oid)
            roles = usr.get("roles") or ["Root"]      # the reference fixtures create super users
            for rname in roles:
                role = db.one("SELECT id FROM role WHERE name=?", (rname,))
                if role:
                    db.execute("INSERT OR IGNORE INTO user_role VALUES (?,?)", (uid, role["id"]))
            counts["users"] += 1
    for col in entities.get("collaborations", []) or []:
        existing = db.one("SELECT id FROM collaboration WHERE name=?", (col["name"],))
        cid = existing["id"] if existing else db.insert("collaboration", name=col["name"],
                                                          encrypted=1 if col.get("encrypted") else 0)
        if not existing:
            counts["collaborations"] += 1
        for part in col.get("participants", []) or []:
            oid = org_ids.get(part["name"])
            if oid is None:
                o = db.one("SELECT id FROM organization WHERE name=?", (part["name"],))
                if o is None:
                    raise ValueError(f"participant {part['name']!r} is not a known organization")
                oid = o["id"]
            db.execute("INSERT OR IGNORE INTO member VALUES (?,?)", (cid, oid))
            if db.one("SELECT id FROM node WHERE organization_id=? AND collaboration_id=?", (oid, cid)) is None:
                key = part.get("api-key") or part.get("api_key") or db.new_api_key()
                db.insert("node", name=f"{part['name']} - {col['name']} Node", api_key=str(key), collaboration_id=cid,
                          organization_id=oid)
                counts["nodes"] += 1
        for task in col.get("tasks", []) or []:
            tid = db.insert("task", name=task.get("name", ""), description=task.get("description", ""),
                            image=task.get("image", ""), collaboration_id=cid, run_id=db.next_run_id(),
                            database=task.get("database", "default"), created_at=now())
            for oid in db.collaboration_organizations(cid):
                inp = task.get("input", {})
                db.insert("result", task_id=tid, organization_id=oid,
                          input=inp if isinstance(inp, str) else json.dumps(inp), assigned_at=now())
            counts["tasks"] += 1
    log.info("imported %s", counts)
    return counts
