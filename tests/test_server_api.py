"""Central server REST API, permissions, events, client library and encryption, in-process
(server on an ephemeral port, real HTTP)."""
import threading
import time

import pytest

from vantage6_b200.client import ContainerClient, ServerError, UserClient
from vantage6_b200.common.encryption import DummyCryptor, RSACryptor
from vantage6_b200.common.serialization import deserialize, serialize
from vantage6_b200.node import NodeClient
from vantage6_b200.server import fixtures
from vantage6_b200.server.app import ServerApp

ENTITIES = {
    "organizations": [
        {"name": "A", "domain": "a.test", "users": [{"username": "alice", "password": "pw-a", "roles": ["Root"]}]},
        {"name": "B", "domain": "b.test", "users": [{"username": "bob", "password": "pw-b", "roles": ["Researcher"]}]},
        {"name": "C", "domain": "c.test", "users": [{"username": "carol", "password": "pw-c", "roles": ["Viewer"]}]},
    ],
    "collaborations": [{"name": "AB", "participants": [{"name": "A", "api-key": "key-a"}, {"name": "B", "api-key": "key-b"}]}],
}


@pytest.fixture
def server():
    app = ServerApp({"uri": "sqlite://", "api_path": "/api", "jwt_secret_key": "s" * 40, "allow_drop_all": True})
    # drop the bootstrap root user/organization so that the fixture ids are A=1, B=2, C=3
    for sql in ("DELETE FROM user_role", "DELETE FROM user", "DELETE FROM organization",
                "DELETE FROM sqlite_sequence WHERE name IN ('organization', 'user')"):
        app.db.execute(sql)
    fixtures.load(app.db, ENTITIES)
    port = app.start("127.0.0.1", 0)
    yield app, port
    app.stop()


def user(port, name, pw):
    c = UserClient("http://127.0.0.1", port, "/api")
    c.authenticate(name, pw)
    c.setup_encryption(None)
    return c


def test_version_health_and_auth_errors(server):
    app, port = server
    c = UserClient("http://127.0.0.1", port, "/api")
    assert c.util.get_server_version()["version"] == "3.1.0"
    assert c.util.get_server_health()["database"] is True
    with pytest.raises(ServerError) as e:
        c.authenticate("alice", "wrong")
    assert e.value.status == 401
    with pytest.raises(ServerError) as e:
        c.request("organization")
    assert e.value.status == 401
    c.authenticate("alice", "pw-a")
    assert c.whoami.organization_name == "A"
    c.refresh_token()
    assert len(c.organization.list()) >= 3


def test_permissions_scopes(server):
    app, port = server
    alice, bob, carol = user(port, "alice", "pw-a"), user(port, "bob", "pw-b"), user(port, "carol", "pw-c")
    assert {o["name"] for o in alice.organization.list()} >= {"A", "B", "C"}
    assert {o["name"] for o in bob.organization.list()} == {"A", "B"}        # collaboration scope
    assert {o["name"] for o in carol.organization.list()} == {"C"}           # organization scope
    with pytest.raises(ServerError) as e:
        bob.organization.create("D")
    assert e.value.status == 401
    new = alice.organization.create("D", domain="d.test")
    assert new["name"] == "D"
    with pytest.raises(ServerError):
        carol.task.create(collaboration=1, organizations=[1], name="x", image="img", input={})
    with pytest.raises(ServerError):
        bob.user.create("eve", "pw", organization=1)
    u = alice.user.create("eve", "pw", organization=new["id"], roles=[r["id"] for r in alice.role.list() if r["name"] == "Viewer"])
    assert u["username"] == "eve"
    assert len(alice.rule.list()) == 9 * 4 * 4
    collab = alice.collaboration.create("AD", [1, new["id"]])
    node = alice.node.create(collab["id"], new["id"])
    assert "api_key" in node
    with pytest.raises(ServerError):
        alice.node.create(collab["id"], new["id"])          # one node per org x collaboration
    alice.node.delete(node["id"])
    alice.collaboration.delete(collab["id"])


def test_task_lifecycle_with_node_and_container_identities(server):
    app, port = server
    bob = user(port, "bob", "pw-b")
    task = bob.task.create(collaboration=1, organizations=[1, 2], name="t", image="v6b200/average",
                           input={"method": "m", "master": True})
    assert not task["complete"] and len(task["results"]) == 2 and task["run_id"] == 1

    node = NodeClient("http://127.0.0.1", port, "/api")
    with pytest.raises(ServerError):
        node.authenticate("wrong-key")
    node.authenticate("key-a", gpu=3)
    assert node.organization_name == "A"
    open_ = node.request("result", params={"state": "open", "node_id": node.node_id, "include": "task"})
    assert len(open_) == 1 and open_[0]["task"]["image"] == "v6b200/average"
    rid = open_[0]["id"]
    inp = deserialize(DummyCryptor().decrypt_str_to_bytes(open_[0]["input"]))
    assert inp == {"method": "m", "master": True}

    # container token: wrong image is refused, correct one can create a sub-task with the same image only
    with pytest.raises(ServerError):
        node.request("token/container", method="post", json={"task_id": task["id"], "image": "other"})
    tok = node.request("token/container", method="post", json={"task_id": task["id"], "image": "v6b200/average"})["container_token"]
    cc = ContainerClient(tok, "http://127.0.0.1", port, "/api")
    sub = cc.create_new_task({"method": "p"}, [1, 2])
    assert sub["parent"]["id"] == task["id"] and sub["run_id"] == task["run_id"]
    assert [o["name"] for o in cc.get_organizations_in_my_collaboration()] == ["A", "B"]
    assert [a["rank"] for a in cc.get_algorithm_addresses(sub["id"])] == [0, 1]

    # address book: the node registers where the algorithm of its sub-task result listens; siblings of the run see it
    sub_rid = [r["id"] for r in node.request("result", params={"task_id": sub["id"]}) if r["organization"]["id"] == 1][0]
    reg = node.request("port", method="post", json={"result_id": sub_rid, "port": 29617, "label": "rendezvous"})
    assert reg["node_id"] == node.node_id and reg["gpu"] == 3
    with pytest.raises(ServerError):                 # another organization's result
        other_rid = [r["id"] for r in node.request("result", params={"task_id": sub["id"]}) if r["organization"]["id"] == 2][0]
        node.request("port", method="post", json={"result_id": other_rid, "port": 1})
    book = cc.get_algorithm_addresses(sub["id"])
    assert book[0]["ports"] == [{"port": 29617, "label": "rendezvous"}] and book[1]["ports"] == []
    assert [p["port"] for p in bob.request("port", params={"run_id": task["run_id"]})] == [29617]
    assert user(port, "carol", "pw-c").request("port", params={"run_id": task["run_id"]}) == []      # not in the collaboration
    node.request("port", method="delete", params={"result_id": sub_rid})
    assert bob.request("port") == []
    spec = {(r["method"], r["path"]) for r in cc.request("spec")}
    assert {("POST", "/api/port"), ("GET", "/api/role/<id>/rule"), ("PATCH", "/api/result/<id>"), ("POST", "/api/recover/lost")} <= spec

    # node reports; a finished result cannot be patched again; other orgs' results are off limits
    node.request(f"result/{rid}", method="patch", json={"started_at": "now"})
    out = DummyCryptor().bytes_to_str(serialize({"answer": 42}))
    node.request(f"result/{rid}", method="patch", json={"finished_at": "now", "result": out, "log": "ok"})
    with pytest.raises(ServerError):
        node.request(f"result/{rid}", method="patch", json={"log": "again"})
    other = [r for r in bob.result.list(task_id=task["id"]) if r["id"] != rid][0]
    with pytest.raises(ServerError):
        node.request(f"result/{other['id']}", method="patch", json={"log": "x"})
    res = bob.result.from_task(task["id"])
    assert res[0]["result"] == {"answer": 42} and res[1]["result"] is None
    assert not bob.task.get(task["id"])["complete"]
    assert node.request(f"node/{node.node_id}")["status"] == "online" and node.request(f"node/{node.node_id}")["gpu"] == 3

    # deleting a task removes its sub-tasks and results
    alice = user(port, "alice", "pw-a")
    alice.task.delete(task["id"])
    assert alice.task.list() == []
    assert app.db.query("SELECT * FROM result") == []


def test_event_long_poll_pushes_new_tasks(server):
    app, port = server
    node = NodeClient("http://127.0.0.1", port, "/api")
    node.authenticate("key-b")
    start = node.request("event", params={"timeout": 0})["last_id"]
    got = {}

    def listen():
        got["reply"] = node.request("event", params={"since": start, "timeout": 10}, timeout=20)

    th = threading.Thread(target=listen)
    th.start()
    time.sleep(0.2)
    bob = user(port, "bob", "pw-b")
    t0 = time.time()
    bob.task.create(collaboration=1, organizations=[2], name="t", image="img", input={})
    th.join(5)
    assert not th.is_alive() and time.time() - t0 < 2.0          # pushed, not polled
    evs = [e for e in got["reply"]["events"] if e["name"] == "new_task"]
    assert evs and evs[0]["data"]["organization_id"] == 2


def test_patch_organization_public_key_and_rsa_roundtrip(server, tmp_path):
    """``vnode create-private-key`` flow + end-to-end encrypted task input (reference node.py:591-614)."""
    app, port = server
    key_b = tmp_path / "b.pem"
    RSACryptor.create_new_rsa_key(key_b, bits=2048)
    cb = RSACryptor(key_b)
    alice = user(port, "alice", "pw-a")
    alice.organization.update(2, public_key=cb.public_key_str)
    assert cb.verify_public_key(alice.organization.get(2)["public_key"])
    key_a = tmp_path / "a.pem"
    RSACryptor.create_new_rsa_key(key_a, bits=2048)
    alice.setup_encryption(str(key_a))
    alice.organization.update(1, public_key=alice.cryptor.public_key_str)
    task = alice.task.create(collaboration=1, organizations=[2], name="enc", image="img", input={"secret": [1, 2, 3]})
    stored = app.db.one("SELECT input FROM result WHERE task_id=?", (task["id"],))["input"]
    assert "secret" not in stored and stored.count("$") == 2            # ciphertext on the server
    assert deserialize(cb.decrypt_str_to_bytes(stored)) == {"secret": [1, 2, 3]}
    # result encrypted for A comes back decrypted through the client
    enc = cb.encrypt_bytes_to_str(serialize({"ok": True}), alice.cryptor.public_key_str)
    rid = app.db.one("SELECT id FROM result WHERE task_id=?", (task["id"],))["id"]
    app.db.update("result", rid, result=enc, finished_at="now")
    assert alice.wait_for_results(task["id"])[0]["result"] == {"ok": True}


def test_fixtures_drop_all_respects_flag():
    app = ServerApp({"uri": "sqlite://", "api_path": "/api", "allow_drop_all": False})
    with pytest.raises(PermissionError):
        fixtures.load(app.db, ENTITIES, drop_all=True)
    counts = fixtures.load(app.db, ENTITIES)
    assert counts["organizations"] == 3 and counts["nodes"] == 2
    assert fixtures.load(app.db, ENTITIES)["organizations"] == 0         # idempotent


def test_silent_nodes_are_marked_offline(server):
    app, port = server
    node = NodeClient("http://127.0.0.1", port, "/api")
    node.authenticate("key-a")
    assert app.db.get("node", node.node_id)["status"] == "online"
    assert app.reap_silent_nodes(timeout_s=3600) == 0
    time.sleep(0.05)
    assert app.reap_silent_nodes(timeout_s=0.01) == 1
    assert app.db.get("node", node.node_id)["status"] == "offline"
    node.request(f"node/{node.node_id}", method="patch", json={"status": "online"})        # next heartbeat revives it
    assert app.db.get("node", node.node_id)["status"] == "online"


def test_database_refuses_non_identifier_columns(tmp_path):
    """Table / column names are the only things interpolated into SQL: they must be known / plain identifiers."""
    import pytest

    from vantage6_b200.server.db import Database

    db = Database(f"sqlite:///{tmp_path}/t.sqlite")
    oid = db.insert("organization", name="o")
    with pytest.raises(ValueError):
        db.update("organization", oid, **{"name=?, domain": "x"})
    with pytest.raises(ValueError):
        db.insert("organization; DROP TABLE user", name="x")
    db.update("organization", oid, name="renamed")
    assert db.get("organization", oid)["name"] == "renamed"
    db.close()


def test_no_privilege_escalation_through_user_roles(server):
    """an Organization Admin must not be able to create or patch a user into a role (or rule) whose
    permissions it does not hold itself."""
    app, port = server
    alice = user(port, "alice", "pw-a")
    roles = {r["name"]: r["id"] for r in alice.role.list()}
    admin = alice.user.create("dan", "pw-d", organization=2, roles=[roles["Organization Admin"]])
    dan = user(port, "dan", "pw-d")
    with pytest.raises(ServerError) as e:
        dan.user.create("mallory", "pw", organization=2, roles=[roles["Root"]])
    assert e.value.status == 401
    with pytest.raises(ServerError) as e:
        dan.request(f"user/{admin['id']}", method="patch", json={"roles": [roles["Root"]]})
    assert e.value.status == 401
    glob = [r["id"] for r in alice.rule.list() if r["scope"] == "global" and r["name"] == "collaboration" and r["operation"] == "delete"]
    with pytest.raises(ServerError) as e:
        dan.request("user", method="post", json={"username": "m2", "password": "pw", "organization_id": 2, "rules": glob})
    assert e.value.status == 401
    # what it does hold, it may hand out
    ok = dan.user.create("erin", "pw-e", organization=2, roles=[roles["Viewer"]])
    assert ok["username"] == "erin"
    # and dan is still no root: deleting another organization's collaboration stays forbidden
    with pytest.raises(ServerError):
        dan.request("collaboration/1", method="delete")


def test_item_endpoints_respect_collaboration_reach(server):
    """GET /task/<id>, /task/<id>/result, /result/<id>, /collaboration/<id>/*, /node/<id> and the
    task event room are only visible inside the collaboration's reach -- not to any authenticated identity."""
    app, port = server
    alice, carol = user(port, "alice", "pw-a"), user(port, "carol", "pw-c")          # carol: Viewer of organization C
    t = alice.task.create(collaboration=1, organizations=[1, 2], name="secret", image="img", input={"method": "m"})
    rid = alice.request("result", params={"task_id": t["id"]})[0]["id"]
    for path in (f"task/{t['id']}", f"task/{t['id']}/result", f"result/{rid}", "collaboration/1/task", "collaboration/1/node",
                 "collaboration/1/organization", "node/1"):
        with pytest.raises(ServerError) as e:
            carol.request(path)
        assert e.value.status == 401, path
        assert alice.request(path) is not None
    with pytest.raises(ServerError) as e:
        carol.request("event", params={"task_id": t["id"], "timeout": 0.1})
    assert e.value.status == 401
    # a node of another collaboration is locked out as well
    new = alice.organization.create("D", domain="d.test")
    cd = alice.collaboration.create("CD", [3, new["id"]])
    n = alice.node.create(cd["id"], organization=3)
    other = NodeClient("http://127.0.0.1", port, "/api")
    other.authenticate(n["api_key"])
    for path in (f"task/{t['id']}", f"result/{rid}", f"task/{t['id']}/result"):
        with pytest.raises(ServerError) as e:
            other.request(path)
        assert e.value.status == 401, path
    # the collaboration's own node still reads its work
    mine = NodeClient("http://127.0.0.1", port, "/api")
    mine.authenticate("key-a")
    assert mine.request(f"task/{t['id']}")["name"] == "secret"


def test_websocket_event_channel_pushes_and_authorises(server):
    """The scalable event channel (server/ws_events.py, one asyncio thread for every listener): events are pushed to the
    rooms an identity may see; a bad token is refused; a Viewer of another organization does not receive the task event."""
    import json as _json

    from websockets.sync.client import connect

    app, port = server
    alice, carol = user(port, "alice", "pw-a"), user(port, "carol", "pw-c")
    ws_port = alice.request("health")["event_port"]
    assert ws_port
    since = alice.request("health")["events"]
    node = NodeClient("http://127.0.0.1", port, "/api")
    node.authenticate("key-a")
    with connect(f"ws://127.0.0.1:{ws_port}/?token={node.token}&since={since}") as ws_node, \
            connect(f"ws://127.0.0.1:{ws_port}/?token={carol.token}&since={since}") as ws_carol:
        t = alice.task.create(collaboration=1, organizations=[1], name="pushed", image="img", input={"method": "m"})
        ev = _json.loads(ws_node.recv(timeout=5))
        while ev["name"] != "new_task":                     # the node's own status change arrives first
            ev = _json.loads(ws_node.recv(timeout=5))
        assert ev["data"]["task_id"] == t["id"]
        with pytest.raises(TimeoutError):
            ws_carol.recv(timeout=0.5)
    # replay after `since`: a listener that connects later still gets the buffered event
    with connect(f"ws://127.0.0.1:{ws_port}/?token={node.token}&since={since}") as ws2:
        ev = _json.loads(ws2.recv(timeout=5))
        while ev["name"] != "new_task":
            ev = _json.loads(ws2.recv(timeout=5))
        assert ev["data"]["task_id"] == t["id"]
    from websockets.exceptions import ConnectionClosed

    with connect(f"ws://127.0.0.1:{ws_port}/?token=garbage") as bad:
        with pytest.raises(ConnectionClosed):
            bad.recv(timeout=5)


def test_role_management_and_its_grant_check(server):
    """Custom roles: an Organization Admin builds roles for its own organization out of rules it holds itself; default
    roles are read-only for it; another organization's roles are invisible; nothing it lacks can be put into a role."""
    app, port = server
    alice = user(port, "alice", "pw-a")
    roles = {r["name"]: r["id"] for r in alice.role.list()}
    alice.user.create("dan", "pw-d", organization=2, roles=[roles["Organization Admin"]])
    dan = user(port, "dan", "pw-d")
    rules = alice.rule.list()
    pick = lambda name, op, scope: next(r["id"] for r in rules if (r["name"], r["operation"], r["scope"]) == (name, op, scope))  # noqa: E731
    assert alice.rule.get(pick("task", "view", "organization"))["name"] == "task"
    role = dan.role.create("Task watcher", "sees the tasks", rules=[pick("task", "view", "organization")])
    assert role["organization_id"] == 2 and len(role["rules"]) == 1
    with pytest.raises(ServerError) as e:            # a rule dan does not hold
        dan.role.create("Sneaky", rules=[pick("collaboration", "delete", "global")])
    assert e.value.status == 401
    with pytest.raises(ServerError) as e:
        dan.role.add_rule(role["id"], pick("task", "view", "global"))
    assert e.value.status == 401
    with pytest.raises(ServerError) as e:            # another organization
        dan.role.create("Elsewhere", organization=1)
    assert e.value.status == 401
    with pytest.raises(ServerError) as e:            # default roles stay as they are
        dan.role.update(roles["Viewer"], name="Hacked")
    assert e.value.status == 401
    dan.role.add_rule(role["id"], pick("result", "view", "organization"))
    assert {r["name"] for r in dan.role.rules(role["id"])} == {"task", "result"}
    dan.role.remove_rule(role["id"], pick("result", "view", "organization"))
    upd = dan.role.update(role["id"], description="only tasks", rules=[pick("task", "view", "organization"), pick("node", "view", "organization")])
    assert upd["description"] == "only tasks" and len(upd["rules"]) == 2
    # the role works: erin holds nothing but it
    dan.user.create("erin", "pw-e", organization=2, roles=[role["id"]])
    erin = user(port, "erin", "pw-e")
    assert erin.task.list() == [] and {n["organization"]["id"] for n in erin.node.list()} == {2}
    with pytest.raises(ServerError):
        erin.user.create("x", "y")
    # bob (organization 2) sees it, carol (organization 3) does not, alice (global) does
    assert role["id"] in {r["id"] for r in user(port, "bob", "pw-b").role.list()}
    carol = user(port, "carol", "pw-c")
    assert role["id"] not in {r["id"] for r in carol.role.list()}
    with pytest.raises(ServerError):
        carol.role.get(role["id"])
    assert alice.role.get(role["id"])["users"][0]["id"] == erin.whoami.id_
    # deletion: refused while assigned, unless the dependents are dropped too; default roles never
    with pytest.raises(ServerError) as e:
        dan.role.delete(role["id"])
    assert e.value.status == 400
    with pytest.raises(ServerError) as e:
        alice.role.delete(roles["Viewer"])
    assert e.value.status == 400
    dan.role.delete(role["id"], delete_dependents=True)
    assert erin.user.get()["roles"] == []
    with pytest.raises(ServerError) as e:
        alice.role.get(role["id"])
    assert e.value.status == 404


def test_password_change_recovery_and_lockout():
    app = ServerApp({"uri": "sqlite://", "api_path": "/api", "jwt_secret_key": "s" * 40,
                     "password_policy": {"max_failed_attempts": 3, "inactivation_minutes": 10, "min_length": 6}})
    port = app.start("127.0.0.1", 0)
    try:
        root = user(port, "root", "root")
        root.user.update(email="root@a.test")
        with pytest.raises(ServerError) as e:
            root.util.change_my_password("wrong", "new-password")
        assert e.value.status == 401
        with pytest.raises(ServerError) as e:
            root.util.change_my_password("root", "short")
        assert e.value.status == 400
        root.util.change_my_password("root", "new-password")
        with pytest.raises(ServerError):
            user(port, "root", "root")
        user(port, "root", "new-password")
        # recovery: same reply whether or not the account exists; the token works once
        anon = UserClient("http://127.0.0.1", port, "/api")
        assert anon.util.reset_my_password(username="nobody") == anon.util.reset_my_password(email="root@a.test")
        assert len(app.outbox) == 1 and app.outbox[0]["username"] == "root"
        token = app.outbox[0]["reset_token"]
        with pytest.raises(ServerError) as e:
            anon.util.set_my_password(token + "x", "another-password")
        assert e.value.status == 401
        access = user(port, "root", "new-password").token
        with pytest.raises(ServerError):                       # an access token is not a reset token
            anon.util.set_my_password(access, "another-password")
        anon.util.set_my_password(token, "another-password")
        user(port, "root", "another-password")
        with pytest.raises(ServerError) as e:                  # bound to the password it was issued for
            anon.util.set_my_password(token, "third-password")
        assert e.value.status == 401
        # lockout after three failures, even with the right password; the recovery path reopens the account
        for _ in range(3):
            with pytest.raises(ServerError):
                user(port, "root", "guess")
        with pytest.raises(ServerError) as e:
            user(port, "root", "another-password")
        assert "blocked" in e.value.msg
        anon.util.reset_my_password(username="root")
        anon.util.set_my_password(app.outbox[-1]["reset_token"], "fourth-password")
        user(port, "root", "fourth-password")
    finally:
        app.stop()


def test_membership_sub_resources(server):
    app, port = server
    alice, bob, carol = user(port, "alice", "pw-a"), user(port, "bob", "pw-b"), user(port, "carol", "pw-c")
    assert [o["id"] for o in alice.collaboration.add_organization(1, 3)] == [1, 2, 3]
    with pytest.raises(ServerError) as e:
        bob.collaboration.add_organization(1, 3)
    assert e.value.status == 401
    assert [c["name"] for c in carol.organization.collaborations()] == ["AB"]
    node = alice.node.create(1, 3)
    assert [n["id"] for n in alice.organization.nodes(3)] == [node["id"]]
    with pytest.raises(ServerError) as e:            # a member with a node cannot be dropped
        alice.collaboration.remove_organization(1, 3)
    assert e.value.status == 400
    # detach: the node keeps its key but cannot log in; attach it again and it can
    assert node["id"] not in [n["id"] for n in alice.collaboration.remove_node(1, node["id"])]
    nc = NodeClient("http://127.0.0.1", port, "/api")
    with pytest.raises(ServerError) as e:
        nc.authenticate(node["api_key"])
    assert e.value.status == 401
    assert alice.node.get(node["id"])["collaboration"] is None
    assert node["id"] in [n["id"] for n in alice.collaboration.add_node(1, node["id"])]
    nc.authenticate(node["api_key"])
    # tasks of a node
    t = alice.task.create(collaboration=1, organizations=[1, 3], name="t", image="img", input={"method": "m"})
    assert [x["id"] for x in alice.node.tasks(node["id"])] == [t["id"]]
    assert [x["id"] for x in alice.node.tasks(node["id"], open_only=True)] == [t["id"]]
    assert [x["id"] for x in nc.request(f"node/{node['id']}/task")] == [t["id"]]
    b_node = next(n for n in alice.node.list() if n["organization"]["id"] == 2)
    assert alice.node.tasks(b_node["id"]) == []
    with pytest.raises(ServerError):                 # a node asks about itself only, unless it shares the collaboration view
        UserClient("http://127.0.0.1", port, "/api").request(f"node/{node['id']}/task")
    alice.node.delete(node["id"])
    assert [o["id"] for o in alice.collaboration.remove_organization(1, 3)] == [1, 2]
    assert carol.organization.collaborations() == []


def test_signing_key_is_kept_with_the_database(tmp_path):
    """Without ``jwt_secret_key`` two processes on one database (a restarted server, ``vserver shell``) share the key:
    an operator-minted reset token is accepted by the running server."""
    from vantage6_b200.server.admin_routes import issue_reset_token

    cfg = {"uri": f"sqlite:///{tmp_path}/s.sqlite", "api_path": "/api"}
    app = ServerApp(cfg)
    port = app.start("127.0.0.1", 0)
    try:
        shell = ServerApp(cfg)                                    # what `vserver shell` builds
        assert shell.secret == app.secret and len(app.secret) == 64
        token = issue_reset_token(shell, "root")
        UserClient("http://127.0.0.1", port, "/api").util.set_my_password(token, "better-password")
        user(port, "root", "better-password")
        with pytest.raises(KeyError):
            issue_reset_token(shell, "nobody")
    finally:
        app.stop()


def test_json_http_transport_reconnects_and_drops_none_params(server):
    from vantage6_b200.common.jsonhttp import JsonHttp

    app, port = server
    h = JsonHttp()
    base = f"http://127.0.0.1:{port}/api"
    assert h.request("GET", base + "/version").json()["version"] == "3.1.0"
    conn = h._idle[("http", "127.0.0.1", port)][0]
    assert h.request("GET", base + "/version", params={"x": None, "y": 1}).status_code == 200
    assert h._idle[("http", "127.0.0.1", port)] == [conn]                 # kept alive and reused
    conn.sock.close()                                                   # the peer (or a firewall) dropped the idle connection
    assert h.request("GET", base + "/version").status_code == 200       # noticed before sending: fresh connection
    assert h._idle[("http", "127.0.0.1", port)][0] is not conn
    r = h.request("POST", base + "/token/user", json={"username": "alice", "password": "nope"})
    assert r.status_code == 401 and "Invalid" in r.json()["msg"]
    assert h.request("GET", base + "/nothing-here").status_code == 404
    results = []
    threads = [threading.Thread(target=lambda: results.append(h.request("GET", base + "/health").status_code)) for _ in range(8)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert results == [200] * 8                                         # a connection serves one request at a time
    assert 1 <= h.idle_connections() <= 8
    done = []
    t = threading.Thread(target=lambda: done.append(h.request("GET", base + "/version").status_code))   # a new thread reuses a warm one
    before = h.idle_connections()
    t.start()
    t.join()
    assert done == [200] and h.idle_connections() == before
    h.close()
    assert h.idle_connections() == 0


def test_fewer_round_trips_per_work_item(server):
    """The node's own room gets the whole work item, the token request reports the start, status events say whether they
    completed the task -- and a client that waits is not woken into useless GETs."""
    app, port = server
    alice = user(port, "alice", "pw-a")
    node = NodeClient("http://127.0.0.1", port, "/api")
    node.authenticate("key-a")
    since = app.events.last_id()
    task = alice.task.create(collaboration=1, organizations=[1, 2], name="t", image="img", input={"method": "m"})
    evs = node.request("event", params={"since": since, "timeout": 1})["events"]
    mine = [e for e in evs if e["name"] == "new_task" and "result" in e["data"]]
    assert len(mine) == 1 and mine[0]["data"]["result"]["task"]["image"] == "img"          # only its own work item, with the task block
    assert mine[0]["data"]["result"]["task"]["initiator"] == 1
    assert sum(e["name"] == "new_task" for e in evs) == 3                                  # + the two collaboration-wide notices
    bob_events = user(port, "bob", "pw-b").request("event", params={"since": since, "timeout": 1})["events"]
    assert all("result" not in e["data"] for e in bob_events)                              # nobody else sees work items
    rid = mine[0]["data"]["result"]["id"]
    reply = node.request("token/container", method="post", json={"task_id": task["id"], "image": "img", "result_id": rid})
    assert reply["started"] is True
    r = alice.result.get(rid)
    assert r["started_at"] is not None and r["status"] == "active"
    other = [x["id"] for x in alice.result.list(task_id=task["id"]) if x["id"] != rid][0]
    assert "started" not in node.request("token/container", method="post", json={"task_id": task["id"], "image": "img", "result_id": other})
    assert alice.result.get(other)["started_at"] is None                                   # not this node's result: untouched
    # a waiting client: woken by the first completion (not complete), returns on the second without extra polling
    calls = []
    real = alice.request
    alice.request = lambda endpoint, *a, **k: (calls.append(endpoint), real(endpoint, *a, **k))[1]
    got = {}
    waiter = threading.Thread(target=lambda: got.update(rows=alice.wait_for_results(task["id"], timeout=20)))
    waiter.start()
    time.sleep(0.3)
    node.request(f"result/{rid}", method="patch", json={"finished_at": "now", "result": DummyCryptor().bytes_to_str(serialize({"v": 1}))})
    time.sleep(0.3)
    node_b = NodeClient("http://127.0.0.1", port, "/api")
    node_b.authenticate("key-b")
    node_b.request(f"result/{other}", method="patch", json={"finished_at": "now", "result": DummyCryptor().bytes_to_str(serialize({"v": 2}))})
    waiter.join(timeout=20)
    assert sorted(r["result"]["v"] for r in got["rows"]) == [1, 2]
    assert calls.count(f"task/{task['id']}") <= 2 and f"task/{task['id']}/result" not in calls   # first check + the completing event
    done = [e for e in node.request("event", params={"since": since, "timeout": 1})["events"] if e["name"] == "status_update"]
    assert [e["data"]["task_complete"] for e in done] == [False, False, True]              # started, first result, last result


def test_metrics_endpoint_prometheus_text(server):
    import requests

    app, port = server
    alice = user(port, "alice", "pw-a")
    alice.task.create(collaboration=1, organizations=[1, 2], name="t", image="img", input={})
    with pytest.raises(ServerError):
        alice.request("task/999")
    url = f"http://127.0.0.1:{port}/api/metrics"
    s = requests.Session()
    s.trust_env = False
    assert s.get(url).status_code == 401                                   # a token is needed unless metrics_public is set
    r = s.get(url, headers=alice.headers)
    assert r.status_code == 200 and r.headers["Content-Type"].startswith("text/plain")
    lines = r.text.splitlines()
    assert 'v6_http_requests_total{method="POST",route="/task",status="201"} 1' in lines
    assert 'v6_http_requests_total{method="GET",route="/task/<id>",status="404"} 1' in lines
    assert 'v6_nodes{status="offline"} 2' in lines and 'v6_results{status="pending"} 2' in lines
    assert "v6_tasks_open 1" in lines and "v6_tasks_total 1" in lines
    assert any(l.startswith('v6_http_request_seconds_total{method="POST",route="/token/user"}') for l in lines)
    for l in lines:                                                        # well-formed exposition
        assert l.startswith("# ") or (" " in l and float(l.rsplit(" ", 1)[1]) >= 0)
    app.config["metrics_public"] = True
    assert s.get(url).status_code == 200


def test_grant_cap_holds_for_random_rule_sets():
    """Property (hypothesis): whatever rules a user holds, it can put a rule into a role exactly when it holds that
    (resource, operation) itself at an equal or wider scope -- checked against an independent oracle, in-process."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    from vantage6_b200.server.app import SCOPES

    app = ServerApp({"uri": "sqlite://", "api_path": "/api", "jwt_secret_key": "s" * 40})
    rules = app.db.query("SELECT * FROM rule ORDER BY id")
    by_id = {r["id"]: r for r in rules}
    org = app.db.one("SELECT id FROM organization WHERE name='root'")["id"]
    counter = [0]

    def call(method, path, body, uid):
        tok = app.make_token("user", {"id": uid, "organization_id": org})
        return app.dispatch(method, "/api" + path, {}, body, {"Authorization": "Bearer " + tok})

    ids = st.lists(st.sampled_from([r["id"] for r in rules]), min_size=0, max_size=12, unique=True)

    @settings(max_examples=80, deadline=None)
    @given(ids, ids)
    def check(held, wanted):
        counter[0] += 1
        uid = app.db.insert("user", username=f"u{counter[0]}", password="x", organization_id=org)
        for rid in held:
            app.db.execute("INSERT INTO user_rule VALUES (?,?)", (uid, rid))
        status, payload = call("POST", "/role", {"name": f"r{counter[0]}", "rules": wanted}, uid)
        best = {}
        for rid in held:
            r = by_id[rid]
            key = (r["name"], r["operation"])
            best[key] = max(best.get(key, -1), SCOPES.index(r["scope"]))
        may_create = ("role", "create") in best
        covered = all(best.get((by_id[w]["name"], by_id[w]["operation"]), -1) >= SCOPES.index(by_id[w]["scope"]) for w in wanted)
        assert (status == 201) == (may_create and covered), (status, payload, held, wanted)
        if status == 201:
            assert sorted(x["id"] for x in payload["rules"]) == sorted(wanted)

    check()


def test_event_log_keeps_only_recent_work_item_payloads():
    from vantage6_b200.server.app import EventBus

    bus = EventBus()
    bus.HEAVY_KEPT = 3
    for i in range(5):
        bus.emit("new_task", {"result_id": i, "result": {"input": "x" * 100}}, ["node_1"])
        bus.emit("new_task", {"result_id": i}, ["collaboration_1"])
    evs = bus.wait(0, ["node_1"], 0.0)
    assert [("result" in e["data"]) for e in evs] == [False, False, True, True, True]       # old payloads dropped, ids kept
    assert [e["data"]["result_id"] for e in evs] == [0, 1, 2, 3, 4]


def _self_signed(tmp_path):
    import datetime
    import ipaddress

    from cryptography import x509
    from cryptography.hazmat.primitives import hashes, serialization
    from cryptography.hazmat.primitives.asymmetric import rsa
    from cryptography.x509.oid import NameOID

    key = rsa.generate_private_key(public_exponent=65537, key_size=2048)
    name = x509.Name([x509.NameAttribute(NameOID.COMMON_NAME, "vantage6-b200 test server")])
    now = datetime.datetime.now(datetime.timezone.utc)
    cert = (x509.CertificateBuilder().subject_name(name).issuer_name(name).public_key(key.public_key())
            .serial_number(x509.random_serial_number()).not_valid_before(now - datetime.timedelta(minutes=5))
            .not_valid_after(now + datetime.timedelta(days=2))
            .add_extension(x509.SubjectAlternativeName([x509.IPAddress(ipaddress.ip_address("127.0.0.1")), x509.DNSName("localhost")]), False)
            .add_extension(x509.BasicConstraints(ca=True, path_length=None), True)
            .sign(key, hashes.SHA256()))
    certfile, keyfile = tmp_path / "server.crt", tmp_path / "server.key"
    certfile.write_bytes(cert.public_bytes(serialization.Encoding.PEM))
    keyfile.write_bytes(key.private_bytes(serialization.Encoding.PEM, serialization.PrivateFormat.TraditionalOpenSSL, serialization.NoEncryption()))
    return str(certfile), str(keyfile)


def test_tls_for_rest_and_event_channel(tmp_path):
    """``ssl: {certfile, keyfile}`` in the server configuration: REST over https and events over wss, verified against the
    certificate the client was given -- and refused without it."""
    import json as _json
    import ssl

    from websockets.sync.client import connect

    certfile, keyfile = _self_signed(tmp_path)
    app = ServerApp({"uri": "sqlite://", "api_path": "/api", "jwt_secret_key": "s" * 40, "ssl": {"certfile": certfile, "keyfile": keyfile}})
    port = app.start("127.0.0.1", 0)
    try:
        c = UserClient("https://127.0.0.1", port, "/api", ca_file=certfile)
        c.authenticate("root", "root")
        c.setup_encryption(None)
        assert c.util.get_server_health()["database"] is True
        for _ in range(3):                                             # kept-alive TLS connection is reused
            assert c.util.get_server_version()["version"] == "3.1.0"
        assert c._http.idle_connections() == 1
        with pytest.raises(ssl.SSLCertVerificationError):              # unknown issuer without the CA file
            UserClient("https://127.0.0.1", port, "/api").util.get_server_version()
        with pytest.raises(Exception):                                 # plain http against the TLS port
            UserClient("http://127.0.0.1", port, "/api").util.get_server_version()
        org = c.organization.create("O")
        collab = c.collaboration.create("C", [org["id"]])
        node = c.node.create(collab["id"], org["id"])
        nc = NodeClient("https://127.0.0.1", port, "/api", ca_file=certfile)
        nc.authenticate(node["api_key"])
        ev_port = c.util.get_server_health()["event_port"]
        with connect(f"wss://127.0.0.1:{ev_port}/?token={nc.token}&since={app.events.last_id()}", ssl=nc._http.ssl_context(), open_timeout=10) as ws:
            # the root user is in organization "root": create the task as a member of the collaboration instead
            c.user.create("res", "pw-res", organization=org["id"], roles=[r["id"] for r in c.role.list() if r["name"] == "Researcher"])
            res = UserClient("https://127.0.0.1", port, "/api", ca_file=certfile)
            res.authenticate("res", "pw-res")
            res.setup_encryption(None)
            task = res.task.create(collaboration=collab["id"], organizations=[org["id"]], name="t", image="img", input={"m": 1})
            ev = _json.loads(ws.recv(timeout=10))
            assert ev["name"] == "new_task" and ev["data"]["task_id"] == task["id"] and "result" in ev["data"]
    finally:
        app.stop()


def test_tls_client_survives_a_server_restart(tmp_path):
    """A kept-alive TLS connection whose server went away is replaced on the next request (also for a POST)."""
    certfile, keyfile = _self_signed(tmp_path)
    cfg = {"uri": f"sqlite:///{tmp_path}/s.sqlite", "api_path": "/api", "jwt_secret_key": "s" * 40, "ssl": {"certfile": certfile, "keyfile": keyfile}}
    app = ServerApp(cfg)
    port = app.start("127.0.0.1", 0)
    c = UserClient("https://127.0.0.1", port, "/api", ca_file=certfile)
    c.authenticate("root", "root")
    assert c._http.idle_connections() == 1
    app.stop()
    app2 = ServerApp(cfg)
    app2.start("127.0.0.1", port)
    try:
        made = c.request("organization", method="post", json={"name": "after-restart"})          # POST on the stale connection
        assert made["name"] == "after-restart" and c.util.get_server_health()["database"] is True
    finally:
        app2.stop()


def test_concurrent_researchers_and_nodes_keep_the_books_straight(tmp_path):
    """Four researchers create tasks while four nodes start and finish the work items, all at once, against a file
    database: no request fails, every result is started and finished exactly once, event ids are unique and ordered."""
    app = ServerApp({"uri": f"sqlite:///{tmp_path}/s.sqlite", "api_path": "/api", "jwt_secret_key": "s" * 40})
    entities = {"organizations": [{"name": f"O{i}", "users": [{"username": f"u{i}", "password": "pw", "roles": ["Researcher"]}]} for i in range(4)],
                "collaborations": [{"name": "C", "participants": [{"name": f"O{i}", "api-key": f"k{i}"} for i in range(4)]}]}
    fixtures.load(app.db, entities)
    port = app.start("127.0.0.1", 0)
    orgs = [o["id"] for o in app.db.query("SELECT id FROM organization WHERE name LIKE 'O%' ORDER BY id")]
    cid = app.db.one("SELECT id FROM collaboration")["id"]
    errors, created, done = [], [], []
    per_researcher = 10

    def researcher(i):
        try:
            c = user(port, f"u{i}", "pw")
            for k in range(per_researcher):
                t = c.task.create(collaboration=cid, organizations=orgs, name=f"t{i}-{k}", image="img", input={"k": k})
                created.append(t["id"])
                c.task.list()
                c.result.list(task_id=t["id"])
        except Exception as e:  # noqa: BLE001
            errors.append(("researcher", i, repr(e)))

    def node(i):
        try:
            n = NodeClient("http://127.0.0.1", port, "/api")
            n.authenticate(f"k{i}")
            seen, deadline = set(), time.time() + 60
            while time.time() < deadline and len(seen) < 4 * per_researcher:
                for r in n.request("result", params={"state": "open", "node_id": n.node_id}):
                    if r["id"] in seen:
                        continue
                    seen.add(r["id"])
                    assert n.request("token/container", method="post", json={"task_id": r["task"]["id"], "image": "img", "result_id": r["id"]})["started"]
                    n.request(f"result/{r['id']}", method="patch", json={"finished_at": "now", "result": "x", "status": "completed"})
                    done.append(r["id"])
                time.sleep(0.01)
        except Exception as e:  # noqa: BLE001
            errors.append(("node", i, repr(e)))

    threads = [threading.Thread(target=researcher, args=(i,)) for i in range(4)] + [threading.Thread(target=node, args=(i,)) for i in range(4)]
    try:
        [t.start() for t in threads]
        [t.join(timeout=90) for t in threads]
        assert not errors, errors[:3]
        n_results = 4 * 4 * per_researcher
        assert len(created) == 4 * per_researcher and sorted(done) == sorted(set(done)) and len(done) == n_results
        books = app.db.one("SELECT COUNT(*) AS n, SUM(finished_at IS NOT NULL) AS f, SUM(started_at IS NOT NULL) AS s FROM result")
        assert (books["n"], books["f"], books["s"]) == (n_results, n_results, n_results)
        evs = app.events.wait(0, [f"collaboration_{cid}"], 0.0)
        ids = [e["id"] for e in evs]
        assert ids == sorted(set(ids))
        names = [e["name"] for e in evs]
        assert names.count("new_task") == n_results and names.count("status_update") == 2 * n_results          # started + finished
        completing = {e["data"]["task_id"] for e in evs if e["name"] == "status_update" and e["data"].get("task_complete")}
        assert completing == set(created)                  # every task announced complete (two racing finishers may both say so)
    finally:
        app.stop()


def test_a_node_cannot_be_taken_from_another_collaboration(server):
    """Attaching an attached node moves it: allowed only to someone who may edit BOTH collaborations."""
    app, port = server
    alice = user(port, "alice", "pw-a")
    roles = {r["name"]: r["id"] for r in alice.role.list()}
    other = alice.collaboration.create("BC", [2, 3])                                  # bob's organization is in AB and in BC
    alice.user.create("cadmin", "pw", organization=3, roles=[roles["Collaboration Admin"]])      # organization 3: only in BC
    cadmin = user(port, "cadmin", "pw")
    b_node = next(n for n in alice.node.list() if n["organization"]["id"] == 2)       # B's node, attached to AB
    with pytest.raises(ServerError) as e:
        cadmin.collaboration.add_node(other["id"], b_node["id"])                       # may edit BC, not AB
    assert e.value.status == 401
    assert alice.node.get(b_node["id"])["collaboration"]["id"] == 1
    moved = alice.collaboration.add_node(other["id"], b_node["id"])                    # global scope: fine
    assert b_node["id"] in [n["id"] for n in moved]


def test_a_node_may_only_publish_its_organizations_key(server):
    app, port = server
    node = NodeClient("http://127.0.0.1", port, "/api")
    node.authenticate("key-a")
    node.request("organization/1", method="patch", json={"public_key": "abc"})
    assert app.db.get("organization", 1)["public_key"] == "abc"
    with pytest.raises(ServerError) as e:
        node.request("organization/1", method="patch", json={"name": "renamed", "public_key": "x"})
    assert e.value.status == 401 and app.db.get("organization", 1)["name"] == "A" and app.db.get("organization", 1)["public_key"] == "abc"
    with pytest.raises(ServerError):
        node.request("organization/2", method="patch", json={"public_key": "x"})           # not its organization


def test_tokens_die_with_their_account(server):
    app, port = server
    alice = user(port, "alice", "pw-a")
    eve = alice.user.create("eve", "pw-e", organization=2, roles=[r["id"] for r in alice.role.list() if r["name"] == "Researcher"])
    eve_client = user(port, "eve", "pw-e")
    assert eve_client.task.list() == []
    alice.user.delete(eve["id"])
    with pytest.raises(ServerError) as e:
        eve_client.task.list()
    assert e.value.status == 401 and "deleted" in e.value.msg
    node = NodeClient("http://127.0.0.1", port, "/api")
    node.authenticate("key-b")
    node.request("result", params={"state": "open", "node_id": node.node_id})
    alice.node.delete(node.node_id)
    with pytest.raises(ServerError) as e:
        node.request("result", params={"state": "open"})
    assert e.value.status == 401


def test_oversized_or_malformed_content_length_is_refused_unread(server):
    import socket as _socket

    app, port = server
    for header, status in ((b"Content-Length: 999999999999", b"413"), (b"Content-Length: banana", b"400"), (b"Content-Length: -5", b"400")):
        with _socket.create_connection(("127.0.0.1", port), timeout=5) as s:
            s.sendall(b"POST /api/token/user HTTP/1.1\r\nHost: x\r\nContent-Type: application/json\r\n" + header + b"\r\n\r\n{}")
            reply = s.recv(4096)
            assert reply.startswith(b"HTTP/1.1 " + status) or reply.startswith(b"HTTP/1.0 " + status), reply[:80]
            assert b"Connection: close" in reply
    assert UserClient("http://127.0.0.1", port, "/api").util.get_server_version()["version"] == "3.1.0"      # the server is fine
