"""Unit tests of the shared managed-instance layer (cli/instance.py): naming, option factories, the launch plan,
the guards and the name prompt -- the pieces both ``vnode`` and ``vserver`` are assembled from."""
from types import SimpleNamespace
from unittest.mock import MagicMock

import click
import pytest
from click.testing import CliRunner

from vantage6_b200.cli import instance
from vantage6_b200.cli.instance import NODE, SERVER, LaunchPlan


def test_runtime_names_and_labels():
    assert NODE.runtime_name("iknl", False) == "vantage6-iknl-user"
    assert NODE.runtime_name("iknl", True) == "vantage6-iknl-system"
    assert SERVER.runtime_name("iknl", True) == "vantage6-iknl-system-server"
    assert (NODE.label, SERVER.label) == ("vantage6-type=node", "vantage6-type=server")
    assert (NODE.plural, SERVER.noun, SERVER.local_cli) == ("nodes", "Server", "vserver-local")


def test_running_filters_by_label():
    rt = MagicMock()
    rt.containers.list.return_value = [SimpleNamespace(name="vantage6-a-user"), SimpleNamespace(name="vantage6-b-user")]
    assert NODE.running(rt) == ["vantage6-a-user", "vantage6-b-user"]
    rt.containers.list.assert_called_once_with(filters={"label": "vantage6-type=node"})


def test_table_header_layout():
    header = instance.table_header()
    assert header.startswith("\nName") and len(header) == 85
    assert header.index("Environments") == 1 + 25 and header.index("Status") == 1 + 25 + 32


def test_launch_plan_collects_binds_in_order():
    plan = LaunchPlan(name="vantage6-x-user", keep=True)
    plan.bind("/mnt/log", "/var/log/x")
    plan.bind("/mnt/data", "vol-x")
    plan.environment.update(A="1")
    assert plan.volume_specs() == ["/var/log/x:/mnt/log", "vol-x:/mnt/data"]
    assert plan.keep and plan.environment == {"A": "1"} and plan.labels == {}


@pytest.mark.parametrize("name,ok", [("iknl", True), ("a.b-c_d9", True), ("bad$name", False), ("with space", False), ("", False)])
def test_name_guard(name, ok):
    if ok:
        instance.check_config_name_allowed(name)
    else:
        with pytest.raises(SystemExit) as e:
            instance.check_config_name_allowed(name)
        assert e.value.code == 1


def test_runtime_ping_guard():
    rt = MagicMock()
    instance.check_if_docker_deamon_is_running(rt)
    rt.ping.side_effect = RuntimeError("no runtime")
    with pytest.raises(SystemExit):
        instance.check_if_docker_deamon_is_running(rt)


def test_name_prompt_normalisation(capsys):
    deps = SimpleNamespace(q=MagicMock())
    deps.q.text.return_value.ask.return_value = "my node"
    assert instance.ask_configuration_name(deps, None, always_normalise=True) == "my-node"
    assert "Replaced spaces from configuration name: my-node" in capsys.readouterr().out
    assert instance.ask_configuration_name(deps, "given name", always_normalise=False) == "given name"   # server: as typed
    assert instance.ask_configuration_name(deps, "given name", always_normalise=True) == "given-name"


def test_option_factories_build_a_working_command():
    @click.command()
    @instance.name_option()
    @instance.environment_option("application")
    @instance.folders_option(False)
    @instance.config_option()
    def cmd(name, environment, system_folders, config):
        click.echo(f"{name}|{environment}|{system_folders}|{config}")

    run = CliRunner().invoke
    assert run(cmd, []).output.strip() == "None|application|False|None"
    assert run(cmd, ["-n", "x", "-e", "dev", "--system", "-c", "/f.yaml"]).output.strip() == "x|dev|True|/f.yaml"
