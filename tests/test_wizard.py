"""The configuration wizard: questionnaires (node, server), the driver that writes / extends the YAML file, and
the configuration selector.  Scenario coverage follows the reference's wizard tests (tests/test_wizard.py) -- with
its stale ``"rabbitmq"`` key assertion corrected to ``"rabbitmq_uri"`` -- plus the
multi-environment merge and the empty-selector error, which the reference leaves untested."""
from contextlib import contextmanager
from pathlib import Path
from unittest.mock import MagicMock, patch

import pytest
import yaml

from vantage6_b200.cli import configuration_wizard as wizard

WIZARD = "vantage6_b200.cli.configuration_wizard"
NODE_KEYS = {"api_key", "server_url", "port", "api_path", "task_dir", "databases", "logging", "encryption", "vpn_subnet"}
SERVER_KEYS = {"description", "ip", "port", "api_path", "uri", "allow_drop_all", "jwt_secret_key", "logging",
               "vpn_server", "rabbitmq_uri"}


def accept_defaults(questions, **_):
    """Answer every prompt with its default; the default database gets a path (it has no default)."""
    return {item["name"]: ("/some/path/db.sqlite" if item["name"] == "default" else item.get("default"))
            for item in questions}


@contextmanager
def scripted_prompts(confirm_answers, prompt=accept_defaults):
    with patch(f"{WIZARD}.q") as q:
        q.prompt.side_effect = prompt
        q.confirm.return_value.ask.side_effect = list(confirm_answers)
        yield q


@pytest.mark.parametrize("kind,confirms,expected_keys", [
    ("node", (True, False, True), NODE_KEYS),          # one extra database, then stop; VPN yes
    ("server", (True, True, True), SERVER_KEYS),       # jwt secret, vpn server, message queue: all yes
])
def test_questionnaires_cover_every_section(kind, confirms, expected_keys):
    with scripted_prompts(confirms):
        if kind == "node":
            config = wizard.node_configuration_questionaire({"data": "/"}, "iknl")
        else:
            config = wizard.server_configuration_questionaire("", "vantage6")
    assert expected_keys <= set(config)
    if kind == "node":
        assert set(config["databases"]) == {"default", "database_1"}
        assert (config["logging"]["file"], config["logging"]["backup_count"]) == ("iknl.log", 5)
    else:
        assert (config["uri"], config["ip"], config["port"]) == ("sqlite:///default.sqlite", "0.0.0.0", "5000")


def test_driver_returns_the_path_it_wrote():
    targets = ("node_configuration_questionaire", "server_configuration_questionaire", "ServerConfigurationManager",
               "NodeConfigurationManager", "NodeContext")
    patches = [patch(f"{WIZARD}.{t}") for t in targets]
    mocks = dict(zip(targets, (p.start() for p in patches)))
    try:
        mocks["NodeContext"].instance_folders.return_value = {"config": "/some/path/"}
        for kind, system_folders in (("node", False), ("server", True)):
            assert wizard.configuration_wizard(kind, "vtg6", "application", system_folders) == Path("/some/path/vtg6.yaml")
    finally:
        for p in patches:
            p.stop()


def test_second_environment_extends_the_same_file(v6home):
    """Running the wizard again for another environment adds to the file instead of replacing it
    (reference configuration_wizard.py:234-242)."""
    fixed = {"api_key": "k", "server_url": "http://localhost", "port": "5000", "api_path": "/api", "task_dir": "/tmp"}

    def answer(questions, **_):
        return {item["name"]: fixed.get(item["name"], item.get("default", "x")) for item in questions}

    with patch(f"{WIZARD}.q") as q:
        q.prompt.side_effect = answer
        q.confirm.return_value.ask.return_value = False
        q.select.return_value.ask.side_effect = ["INFO", "false", "DEBUG", "false"]
        written = [wizard.configuration_wizard("node", "n1", env, False) for env in ("application", "dev")]
    assert written[0] == written[1]
    document = yaml.safe_load(Path(written[0]).read_text())
    assert document["application"]["port"] == 5000                       # coerced by the schema
    assert document["environments"]["dev"]["logging"]["level"] == "DEBUG"
    assert document["environments"]["prod"] == {}


def _one_configuration(name="vtg6", environments=("application",)):
    cfg = MagicMock(available_environments=list(environments))
    cfg.name = name
    return [[cfg], []]


def test_selector_returns_the_picked_pair():
    with patch(f"{WIZARD}.NodeContext") as node_c, patch(f"{WIZARD}.ServerContext") as server_c, \
            patch(f"{WIZARD}.q") as q:
        node_c.available_configurations.return_value = _one_configuration()
        server_c.available_configurations.return_value = _one_configuration()
        q.select.return_value.ask.return_value = ["vtg6", "application"]
        assert tuple(wizard.select_configuration_questionaire("node", True)) == ("vtg6", "application")


def test_selector_without_configurations_raises():
    with patch(f"{WIZARD}.NodeContext") as node_c:
        node_c.available_configurations.return_value = [[], []]
        with pytest.raises(Exception, match="No configurations could be found!"):
            wizard.select_configuration_questionaire("node", False)
