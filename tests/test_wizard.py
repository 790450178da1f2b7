"""Configuration wizard tests (cases mirror reference tests/test_wizard.py; the stale
``"rabbitmq"`` key assertion there is corrected to ``"rabbitmq_uri"``)."""
from pathlib import Path
from unittest.mock import MagicMock, patch

import yaml

from vantage6_b200.cli.configuration_wizard import (configuration_wizard, node_configuration_questionaire,
                                                    select_configuration_questionaire,
                                                    server_configuration_questionaire)

module_path = "vantage6_b200.cli.configuration_wizard"


def prompts(*args, **kwargs):
    result = {}
    for arg in args[0]:
        name = arg["name"]
        if name == "default":               # default db path
            result[name] = "/some/path/db.sqlite"
        else:
            result[name] = arg.get("default")
    return result


def test_node_wizard():
    with patch(f"{module_path}.q") as q:
        q.prompt.side_effect = prompts
        q.confirm.return_value.ask.side_effect = [True, False, True]
        config = node_configuration_questionaire({"data": "/"}, "iknl")
    for key in ["api_key", "server_url", "port", "api_path", "task_dir", "databases", "logging", "encryption",
                "vpn_subnet"]:
        assert key in config
    assert set(config["databases"]) == {"default", "database_1"}
    assert config["logging"]["file"] == "iknl.log" and config["logging"]["backup_count"] == 5


def test_server_wizard():
    with patch(f"{module_path}.q") as q:
        q.prompt.side_effect = prompts
        q.confirm.return_value.ask.side_effect = [True, True, True]
        config = server_configuration_questionaire("", "vantage6")
    for key in ["description", "ip", "port", "api_path", "uri", "allow_drop_all", "jwt_secret_key", "logging",
                "vpn_server", "rabbitmq_uri"]:
        assert key in config
    assert config["uri"] == "sqlite:///default.sqlite" and config["ip"] == "0.0.0.0" and config["port"] == "5000"


@patch(f"{module_path}.node_configuration_questionaire")
@patch(f"{module_path}.server_configuration_questionaire")
@patch(f"{module_path}.ServerConfigurationManager")
@patch(f"{module_path}.NodeConfigurationManager")
@patch(f"{module_path}.NodeContext")
def test_configuration_wizard_interface(context, node_m, server_m, server_q, node_q):
    context.instance_folders.return_value = {"config": "/some/path/"}
    assert configuration_wizard("node", "vtg6", "application", False) == Path("/some/path/vtg6.yaml")
    assert configuration_wizard("server", "vtg6", "application", True) == Path("/some/path/vtg6.yaml")


def test_wizard_writes_and_merges_environments(v6home):
    """A second run for another environment extends the same file (reference wizard :234-242)."""
    answers = {"api_key": "k", "server_url": "http://localhost", "port": "5000", "api_path": "/api", "task_dir": "/tmp"}
    with patch(f"{module_path}.q") as q:
        q.prompt.side_effect = lambda qs, **_: {d["name"]: answers.get(d["name"], d.get("default", "x")) for d in qs}
        q.confirm.return_value.ask.return_value = False
        q.select.return_value.ask.side_effect = ["INFO", "false", "DEBUG", "false"]
        f1 = configuration_wizard("node", "n1", "application", False)
        f2 = configuration_wizard("node", "n1", "dev", False)
    assert f1 == f2
    doc = yaml.safe_load(Path(f1).read_text())
    assert doc["application"]["port"] == 5000                       # coerced by the schema
    assert doc["environments"]["dev"]["logging"]["level"] == "DEBUG"
    assert doc["environments"]["prod"] == {}


@patch(f"{module_path}.NodeContext")
@patch(f"{module_path}.ServerContext")
def test_select_configuration(server_c, node_c):
    config = MagicMock()
    config.name = "vtg6"
    config.available_environments = ["application"]
    server_c.available_configurations.return_value = [[config], []]
    node_c.available_configurations.return_value = [[config], []]
    with patch(f"{module_path}.q") as q:
        q.select.return_value.ask.return_value = ["vtg6", "application"]
        name, env = select_configuration_questionaire("node", True)
    assert (name, env) == ("vtg6", "application")


@patch(f"{module_path}.NodeContext")
def test_select_configuration_empty_raises(node_c):
    node_c.available_configurations.return_value = [[], []]
    try:
        select_configuration_questionaire("node", False)
    except Exception as e:  # noqa: BLE001
        assert str(e) == "No configurations could be found!"
    else:
        raise AssertionError("expected an exception")
