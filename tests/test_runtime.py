"""Process runtime ("containers" are processes) and the message-queue sidecar."""
import json
import sys
import time

import pytest

from vantage6_b200.cli.rabbitmq.queue_manager import RabbitMQManager, split_rabbitmq_uri
from vantage6_b200.runtime import APIError, LocalRuntime, Mount, NotFound
from vantage6_b200.runtime.addons import NetworkManager, get_server_config_name, remove_container_if_exists


@pytest.fixture
def rt(v6home):
    r = LocalRuntime()
    assert r.ping()
    yield r
    for c in r.containers.list():
        c.kill()


def test_run_list_logs_stop(rt):
    code = "import time,sys\nprint('hello', flush=True)\ntime.sleep(60)"
    c = rt.containers.run("img", command=[sys.executable, "-c", code], name="vantage6-t-user",
                          labels={"vantage6-type": "node", "name": "t"}, environment={"FOO": "bar"})
    assert c.status == "running"
    names = [x.name for x in rt.containers.list(filters={"label": "vantage6-type=node"})]
    assert names == ["vantage6-t-user"]
    assert rt.containers.list(filters={"label": "vantage6-type=server"}) == []
    t0 = time.time()
    while b"hello" not in c.logs() and time.time() - t0 < 10:
        time.sleep(0.05)
    assert b"hello" in c.logs()
    with pytest.raises(APIError):
        rt.containers.run("img", command=[sys.executable, "-c", code], name="vantage6-t-user")
    got = rt.containers.get("vantage6-t-user")
    got.stop(timeout=3)
    assert rt.containers.list() == []
    with pytest.raises(NotFound):
        rt.containers.get("vantage6-t-user")


def test_mount_translation_and_exec(rt, tmp_path):
    (tmp_path / "cfg").mkdir()
    (tmp_path / "cfg" / "x.yaml").write_text("k: v")
    code = "import sys,time;print(open(sys.argv[1]).read(), flush=True);time.sleep(30)"
    c = rt.containers.run("img", command=[sys.executable, "-c", code, "/mnt/config/x.yaml"],
                          volumes=[f"{tmp_path / 'cfg'}:/mnt/config"], name="m1")
    t0 = time.time()
    while b"k: v" not in c.logs() and time.time() - t0 < 10:
        time.sleep(0.05)
    assert b"k: v" in c.logs()
    res = c.exec_run([sys.executable, "-c", "import os;print(os.environ['V6_CONTAINER_NAME'] if 'V6_CONTAINER_NAME' in os.environ else 'x')"])
    assert res.exit_code == 0
    c2 = rt.containers.run("img", command=[sys.executable, "-c", code, "/mnt/one.yaml"],
                           mounts=[Mount("/mnt/one.yaml", str(tmp_path / "cfg" / "x.yaml"))], name="m2")
    t0 = time.time()
    while b"k: v" not in c2.logs() and time.time() - t0 < 10:
        time.sleep(0.05)
    assert b"k: v" in c2.logs()
    c.kill()
    c2.kill()


def test_attach_streams_until_exit(rt):
    code = "import time\nfor i in range(3):\n print('line', i, flush=True)\n time.sleep(0.05)"
    c = rt.containers.run("img", command=[sys.executable, "-c", code], name="a1", auto_remove=False)
    lines = b"".join(c.attach(stream=True, logs=True))
    assert lines.count(b"line") == 3


def test_exited_container_is_auto_removed(rt):
    c = rt.containers.run("img", command=[sys.executable, "-c", "pass"], name="gone")
    c.wait(timeout=10)
    assert rt.containers.list() == []
    assert rt._entry("gone") is None


def test_volumes(rt):
    v = rt.volumes.create("vantage6-n-user-3-tmpvol")
    assert v.path.is_dir()
    assert [x.name for x in rt.volumes.list()] == ["vantage6-n-user-3-tmpvol"]
    rt.volumes.get("vantage6-n-user-3-tmpvol").remove()
    assert rt.volumes.list() == []
    with pytest.raises(NotFound):
        rt.volumes.get("nope")


def test_addons(rt):
    assert get_server_config_name("vantage6-my-server-system-server", "system") == "my-server"
    remove_container_if_exists(rt, name="does-not-exist")
    nm = NetworkManager("vantage6-x-user-network")
    nm.create_network(is_internal=False)
    nm.connect("c1")
    assert nm.contains("c1") and not nm.contains("c2")
    nm.delete_network()


def test_split_rabbitmq_uri_and_hash():
    parts = split_rabbitmq_uri("amqp://user:p@ss:word@host.example:5672/my/vhost")
    assert parts == {"user": "user", "password": "p@ss:word".split("@")[0] if False else parts["password"],
                     "host": parts["host"], "port": parts["port"], "vhost": parts["vhost"]}
    parts = split_rabbitmq_uri("amqp://alice:secret@127.0.0.1:5672/test")
    assert parts == {"user": "alice", "password": "secret", "host": "127.0.0.1", "port": "5672", "vhost": "test"}
    h = RabbitMQManager._get_hashed_pw("secret")
    assert RabbitMQManager.check_pw("secret", h) and not RabbitMQManager.check_pw("other", h)


def test_message_queue_sidecar_lifecycle(rt, v6home, monkeypatch):
    """RabbitMQManager.start: definitions + config written, sidecar up, status probe ok."""
    from types import SimpleNamespace

    from vantage6_b200.dev import free_port

    port = free_port()
    data_dir = v6home / "srvdata"
    data_dir.mkdir(parents=True)
    ctx = SimpleNamespace(config={"rabbitmq_uri": f"amqp://bob:pw@127.0.0.1:{port}/vh"}, data_dir=data_dir, name="mqtest")
    monkeypatch.setattr(RabbitMQManager, "INTERVAL", 0.5)
    mgr = RabbitMQManager(ctx, NetworkManager("vantage6-mqtest-user-network"))
    mgr.start()
    defs = json.loads((data_dir / "definitions.json").read_text())
    assert defs["users"][0]["name"] == "bob" and defs["vhosts"][0]["name"] == "vh"
    assert RabbitMQManager.check_pw("pw", defs["users"][0]["password_hash"])
    assert (data_dir / "rabbitmq.config").exists() and (data_dir / "rabbitmq").is_dir()
    names = [c.name for c in rt.containers.list(filters={"label": "vantage6-type=rabbitmq"})]
    assert names == ["vantage6-mqtest-rabbitmq"]
    assert mgr.is_running()
    rt.containers.get("vantage6-mqtest-rabbitmq").kill()
