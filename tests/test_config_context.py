"""Configuration manager, schema, contexts, utils, version (components the reference leaves
untested: context.py, configuration_manager.py -- SURVEY.md section 4)."""
from pathlib import Path

import pytest
import yaml

from vantage6_b200 import __version__
from vantage6_b200.cli.configuration_manager import NodeConfiguration, NodeConfigurationManager, ServerConfiguration
from vantage6_b200.cli.context import NodeContext, ServerContext
from vantage6_b200.cli.utils import check_config_name_allowed
from vantage6_b200.common.schema import And, Optional, Or, Schema, SchemaError, Use

LOGGING = {"level": "DEBUG", "file": "n.log", "use_console": False, "backup_count": 5, "max_size": 1024,
           "format": "%(message)s", "datefmt": "%H:%M"}
NODE = {"api_key": "abc", "server_url": "http://localhost", "port": 5000, "task_dir": "/tmp", "api_path": "/api",
        "databases": {"default": "/data/a.csv"}, "logging": LOGGING, "encryption": {"enabled": False, "private_key": ""}}
SERVER = {"description": "d", "ip": "0.0.0.0", "port": 5000, "api_path": "/api", "uri": "sqlite:///default.sqlite",
          "allow_drop_all": True, "logging": LOGGING}


def test_version_string():
    assert __version__ == "3.1.0"


def test_schema_primitives():
    s = Schema({"a": Use(int), "b": And(Use(str), len), Optional("c"): Or(int, None), "d": {Use(str): Use(str)}})
    out = s.validate({"a": "5", "b": "x", "d": {1: 2}, "extra": True})
    assert out["a"] == 5 and out["d"] == {"1": "2"} and out["extra"] is True
    with pytest.raises(SchemaError):
        s.validate({"a": "x", "b": "y", "d": {}})
    with pytest.raises(SchemaError):
        s.validate({"a": 1, "b": "", "d": {}})
    with pytest.raises(SchemaError):
        s.validate({"a": 1, "d": {}})


def test_node_and_server_schema():
    assert NodeConfiguration(NODE).is_valid
    assert ServerConfiguration(SERVER).is_valid
    bad = dict(NODE, api_key="")
    assert not NodeConfiguration(bad).is_valid
    bad = dict(SERVER, logging=dict(LOGGING, max_size=8))
    assert not ServerConfiguration(bad).is_valid
    bad = dict(SERVER, logging=dict(LOGGING, level="LOUD"))
    assert not ServerConfiguration(bad).is_valid
    assert NodeConfiguration(dict(NODE, port=None)).is_valid        # port may be None


def test_manager_roundtrip_multi_environment(tmp_path):
    m = NodeConfigurationManager("n1")
    m.put("application", NODE)
    m.put("dev", dict(NODE, api_key="dev-key"))
    f = tmp_path / "n1.yaml"
    m.save(f)
    doc = yaml.safe_load(f.read_text())
    assert set(doc) == {"application", "environments"} and set(doc["environments"]) == {"prod", "acc", "test", "dev"}
    m2 = NodeConfigurationManager.from_file(f)
    assert m2.name == "n1" and m2.available_environments == ["application", "dev"]
    assert m2.get("dev")["api_key"] == "dev-key"
    assert m2.has_application and m2.has_environments and not m2.is_empty
    with pytest.raises(SchemaError):
        m2.put("prod", {"api_key": ""})


def _write(home: Path, kind: str, name: str, env: str, cfg: dict, scope="user"):
    d = home / scope / "config" / kind
    d.mkdir(parents=True, exist_ok=True)
    doc = {"application": cfg if env == "application" else {},
           "environments": {e: (cfg if e == env else {}) for e in ("prod", "acc", "test", "dev")}}
    (d / f"{name}.yaml").write_text(yaml.safe_dump(doc))
    return d / f"{name}.yaml"


def test_node_context_names_and_env_overrides(v6home, monkeypatch):
    _write(v6home, "node", "iknl", "application", NODE)
    NodeContext.LOGGING_ENABLED = False
    assert NodeContext.config_exists("iknl", "application", False)
    assert not NodeContext.config_exists("iknl", "prod", False)
    assert not NodeContext.config_exists("nope", "application", False)
    ctx = NodeContext("iknl", "application", False)
    assert ctx.docker_container_name == "vantage6-iknl-user"
    assert ctx.docker_network_name == "vantage6-iknl-user-net"
    assert ctx.docker_volume_name == "vantage6-iknl-user-vol"
    assert ctx.docker_vpn_volume_name == "vantage6-iknl-user-vpn-vol"
    assert ctx.docker_temporary_volume_name(7) == "vantage6-iknl-user-7-tmpvol"
    assert ctx.databases == {"default": "/data/a.csv"} and ctx.get_database_uri() == "/data/a.csv"
    assert ctx.config_file_name == "iknl" and ctx.scope == "user"
    assert ctx.get_data_file("/abs/key.pem") == "/abs/key.pem"
    assert ctx.get_data_file("key.pem") == str(ctx.data_dir / "key.pem")
    monkeypatch.setenv("DATA_VOLUME_NAME", "my-vol")
    monkeypatch.setenv("VPN_VOLUME_NAME", "my-vpn")
    assert ctx.docker_volume_name == "my-vol" and ctx.docker_vpn_volume_name == "my-vpn"
    configs, failed = NodeContext.available_configurations(False)
    assert [c.name for c in configs] == ["iknl"] and failed == []


def test_available_configurations_reports_failed_imports(v6home):
    _write(v6home, "node", "good", "application", NODE)
    bad = v6home / "user" / "config" / "node" / "bad.yaml"
    bad.write_text(yaml.safe_dump({"application": {"api_key": ""}, "environments": {}}))
    configs, failed = NodeContext.available_configurations(False)
    assert [c.name for c in configs] == ["good"] and [Path(f).name for f in failed] == ["bad.yaml"]


def test_server_context_database_uri(v6home, monkeypatch):
    _write(v6home, "server", "srv", "prod", SERVER, scope="system")
    ServerContext.LOGGING_ENABLED = False
    ctx = ServerContext("srv", "prod", True)
    assert ctx.docker_container_name == "vantage6-srv-system-server"
    assert ctx.get_database_uri() == f"sqlite:///{ctx.data_dir / 'default.sqlite'}"     # relative -> data_dir
    monkeypatch.setenv("VANTAGE6_DB_URI", "sqlite:////abs/db.sqlite")
    assert ctx.get_database_uri() == "sqlite:////abs/db.sqlite"
    monkeypatch.setenv("VANTAGE6_CONFIG_NAME", "renamed")
    ext = ServerContext.from_external_config_file(ctx.config_file, "prod", True)
    assert ext.name == "renamed"


def test_logging_setup_writes_rotating_file(v6home):
    _write(v6home, "node", "logme", "application", NODE)
    NodeContext.LOGGING_ENABLED = True
    try:
        ctx = NodeContext("logme", "application", False)
        assert ctx.log_file.name == "logme-user.log"
        ctx.log.info("hello from the test")
        import logging

        for h in logging.getLogger().handlers:
            h.flush()
        assert "hello from the test" in ctx.log_file.read_text()
    finally:
        NodeContext.LOGGING_ENABLED = False
        import logging

        root = logging.getLogger()
        for h in list(root.handlers):
            if getattr(h, "_v6b200", False):
                root.removeHandler(h)
                h.close()


def test_check_config_name_allowed(capsys):
    check_config_name_allowed("good-name_1.2")
    with pytest.raises(SystemExit) as e:
        check_config_name_allowed("bad name")
    assert e.value.code == 1
    assert "[error]" in capsys.readouterr().out


def test_small_pure_functions_properties(tmp_path):
    """hypothesis over the pure helpers: PEP 440 strings, queue URIs, the hybrid encryption envelope."""
    import re

    from hypothesis import given, settings
    from hypothesis import strategies as st

    from vantage6_b200._version import pep440
    from vantage6_b200.cli.rabbitmq.queue_manager import split_rabbitmq_uri
    from vantage6_b200.common.encryption import DummyCryptor, RSACryptor

    @given(st.integers(0, 99), st.integers(0, 99), st.integers(0, 99), st.sampled_from(["alpha", "beta", "candidate", "final"]),
           st.integers(0, 999), st.integers(0, 99))
    def versions(major, minor, patch, stage, build, post):
        v = pep440((major, minor, patch, stage, build, post))
        assert re.fullmatch(r"\d+\.\d+\.\d+(\.(a|b|rc)\d+)?(\.post\d+)?", v)
        assert (stage == "final") == (re.search(r"\.(a|b|rc)\d+", v) is None) and (post > 0) == (".post" in v)

    ident = st.text(alphabet="abcdefghijklmnopqrstuvwxyz0123456789-_.", min_size=1, max_size=12)

    @given(ident, st.text(alphabet="abcXYZ019:!#%^&*()-_+=", min_size=1, max_size=16), ident, st.integers(1, 65535), ident)
    def uris(user, password, host, port, vhost):
        parts = split_rabbitmq_uri(f"amqp://{user}:{password}@{host}:{port}/{vhost}")
        assert parts == {"user": user, "password": password, "host": host, "port": str(port), "vhost": vhost}

    key_a, key_b = tmp_path / "a.pem", tmp_path / "b.pem"
    RSACryptor.create_new_rsa_key(key_a, bits=2048)
    RSACryptor.create_new_rsa_key(key_b, bits=2048)
    a, b = RSACryptor(key_a), RSACryptor(key_b)

    @settings(max_examples=40, deadline=None)
    @given(st.binary(max_size=5000))
    def envelopes(blob):
        sealed = a.encrypt_bytes_to_str(blob, b.public_key_str)
        assert sealed.count("$") == 2 and b.decrypt_str_to_bytes(sealed) == blob
        if blob:
            assert a.encrypt_bytes_to_str(blob, b.public_key_str) != sealed             # fresh key and iv every time
        assert DummyCryptor().decrypt_str_to_bytes(DummyCryptor().encrypt_bytes_to_str(blob, "")) == blob

    versions()
    uris()
    envelopes()
    with pytest.raises(Exception):                                                      # sealed for b: a cannot open it
        a.decrypt_str_to_bytes(a.encrypt_bytes_to_str(b"secret", b.public_key_str))
    assert b.verify_public_key(b.public_key_str) and not b.verify_public_key(a.public_key_str)
