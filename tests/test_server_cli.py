"""vserver CLI unit tests (CliRunner + mocks; cases mirror reference tests/test_server_cli.py and
add shell / version / stop-with-sidecar which the reference leaves untested)."""
from pathlib import Path
from unittest.mock import MagicMock, patch

from click.testing import CliRunner

from vantage6_b200.cli.globals import APPNAME
from vantage6_b200.cli.server import (cli_server_attach, cli_server_configuration_list, cli_server_files,
                                      cli_server_import, cli_server_new, cli_server_shell, cli_server_start,
                                      cli_server_stop, cli_server_version)

RT = "vantage6_b200.runtime.LocalRuntime"
MOD = "vantage6_b200.cli.server"


def container(name):
    c = MagicMock()
    c.name = name
    return c


def ctx_mock(tmp="."):
    ctx = MagicMock(config_file="/file.yaml", data_dir=Path(tmp), scope="system", environment="prod",
                    config_file_name="not-running", docker_container_name=f"{APPNAME}-not-running-system-server")
    ctx.name = "not-running"
    ctx.config = {"uri": "sqlite:///file.db", "port": 9999, "ip": "127.0.0.1"}
    return ctx


@patch(f"{MOD}.RabbitMQManager")
@patch(f"{MOD}.NetworkManager")
@patch(f"{MOD}.pull_if_newer")
@patch(f"{MOD}.ServerContext")
@patch(f"{RT}.containers")
@patch(f"{RT}.ping")
def test_start(ping, containers, context, pull, network, rabbit):
    containers.list.return_value = [container("does-not-matter")]
    containers.run.return_value = True
    context.config_exists.return_value = True
    runner = CliRunner()
    with runner.isolated_filesystem() as d:
        context.return_value = ctx_mock(d)
        result = runner.invoke(cli_server_start, ["--name", "not-running"])
    assert result.exit_code == 0, result.output
    kw = containers.run.call_args.kwargs
    assert kw["command"].startswith("uwsgi --http :5000 --gevent 1000 --http-websockets")
    assert kw["command"].endswith("--pyargv /mnt/config.yaml")
    assert kw["labels"] == {f"{APPNAME}-type": "server", "name": "not-running"}
    assert kw["ports"] == {"5000/tcp": ("127.0.0.1", "9999")}
    assert kw["environment"]["VANTAGE6_DB_URI"].startswith("sqlite:////")
    assert "Message queue disabled!" in result.output
    rabbit.assert_not_called()


@patch(f"{MOD}.RabbitMQManager")
@patch(f"{MOD}.NetworkManager")
@patch(f"{MOD}.pull_if_newer")
@patch(f"{MOD}.ServerContext")
@patch(f"{RT}.containers")
@patch(f"{RT}.ping")
def test_start_with_message_queue(ping, containers, context, pull, network, rabbit):
    containers.list.return_value = []
    context.config_exists.return_value = True
    runner = CliRunner()
    with runner.isolated_filesystem() as d:
        c = ctx_mock(d)
        c.config["rabbitmq_uri"] = "amqp://u:p@127.0.0.1:5672/vh"
        context.return_value = c
        result = runner.invoke(cli_server_start, ["--name", "not-running"])
    assert result.exit_code == 0, result.output
    rabbit.return_value.start.assert_called_once()


@patch(f"{MOD}.ServerContext")
@patch(f"{RT}.containers")
@patch(f"{RT}.ping")
def test_start_already_running(ping, containers, context):
    containers.list.return_value = [container(f"{APPNAME}-not-running-system-server")]
    context.config_exists.return_value = True
    context.return_value = ctx_mock()
    result = CliRunner().invoke(cli_server_start, ["--name", "not-running"])
    assert "already running" in result.output
    assert result.exit_code == 1


@patch(f"{MOD}.ServerContext")
def test_unknown_config_exits(context):
    context.config_exists.return_value = False
    result = CliRunner().invoke(cli_server_files, ["--name", "nope"])
    assert result.output.startswith("[error]")
    assert result.exit_code == 1


@patch(f"{MOD}.ServerContext")
@patch(f"{RT}.containers")
@patch(f"{RT}.ping")
def test_configuration_list(ping, containers, context):
    containers.list.return_value = [container(f"{APPNAME}-iknl-system-server")]
    cfg = MagicMock(available_environments=["prod"])
    cfg.name = "iknl"
    context.available_configurations.return_value = ([cfg], [])
    result = CliRunner().invoke(cli_server_configuration_list)
    assert result.exit_code == 0
    lines = result.output.splitlines()
    assert lines[1].startswith("Name") and "Online" in lines[3] and "Offline" in lines[4]
    assert lines[3].endswith(" System ") and lines[4].endswith(" User   ")


@patch(f"{MOD}.ServerContext")
def test_files(context):
    context.config_exists.return_value = True
    context.return_value = MagicMock(log_file="/log_file.log", config_file="/iknl.yaml")
    context.return_value.get_database_uri.return_value = "sqlite:////db.sqlite"
    result = CliRunner().invoke(cli_server_files, ["--name", "iknl"])
    assert result.exit_code == 0
    assert "Database           = sqlite:////db.sqlite" in result.output


@patch(f"{RT}.images")
@patch(f"{RT}.containers")
@patch(f"{MOD}.print_log_worker")
@patch(f"{MOD}.ServerContext")
@patch(f"{RT}.ping")
def test_import(ping, context, log_worker, containers, images):
    context.config_exists.return_value = True
    runner = CliRunner()
    with runner.isolated_filesystem() as d:
        context.return_value = ctx_mock(d)
        Path("some.yaml").write_text("does-not-matter")
        result = runner.invoke(cli_server_import, ["--name", "iknl", "--drop-all", "some.yaml"])
    assert result.exit_code == 0, result.output
    cmd = containers.run.call_args.kwargs["command"]
    assert cmd.startswith("vserver-local import -c /mnt/config.yaml -e prod --drop-all /mnt/import.yaml")


@patch(f"{MOD}.configuration_wizard")
@patch(f"{MOD}.check_config_write_permissions")
@patch(f"{MOD}.ServerContext")
def test_new(context, permissions, wizard):
    context.config_exists.return_value = False
    permissions.return_value = True
    wizard.return_value = "/some/file.yaml"
    result = CliRunner().invoke(cli_server_new, ["--name", "iknl"])
    assert result.exit_code == 0
    assert "New configuration created" in result.output


@patch(f"{MOD}.ServerContext")
def test_new_existing(context):
    context.config_exists.return_value = True
    result = CliRunner().invoke(cli_server_new, ["--name", "iknl"])
    assert result.exit_code == 1 and "already exists" in result.output


@patch(f"{MOD}.remove_container_if_exists")
@patch(f"{RT}.containers")
@patch(f"{RT}.ping")
def test_stop_kills_server_and_sidecar(ping, containers, remove):
    containers.list.return_value = [container(f"{APPNAME}-iknl-system-server")]
    result = CliRunner().invoke(cli_server_stop, ["--name", "iknl"])
    assert result.exit_code == 0
    containers.get.return_value.kill.assert_called_once()
    assert remove.call_args.kwargs["name"] == f"{APPNAME}-iknl-rabbitmq"
    assert "Stopped the vantage6-iknl-system-server server." in result.output


@patch(f"{RT}.containers")
@patch(f"{RT}.ping")
def test_stop_nothing_running(ping, containers):
    containers.list.return_value = []
    result = CliRunner().invoke(cli_server_stop, ["--name", "iknl"])
    assert result.output.startswith("[warn]")


@patch(f"{MOD}.time.sleep")
@patch(f"{RT}.containers")
@patch(f"{RT}.ping")
def test_attach(ping, containers, sleep):
    containers.list.return_value = [container(f"{APPNAME}-iknl-system-server")]
    containers.get.return_value.attach.return_value = iter([b"a\n"])
    sleep.side_effect = KeyboardInterrupt("Boom!")
    result = CliRunner().invoke(cli_server_attach, ["--name", "iknl"])
    assert result.output.endswith("[info]  - Closing log file. Keyboard Interrupt.\n")
    assert result.exit_code == 0


@patch(f"{RT}.containers")
@patch(f"{RT}.ping")
def test_version_named_lookup_uses_server_suffix(ping, containers):
    """The reference forgets the ``-server`` suffix here; we do not."""
    containers.list.return_value = [container(f"{APPNAME}-iknl-system-server")]
    containers.get.return_value.exec_run.return_value = MagicMock(output=b"3.1.0\n")
    result = CliRunner().invoke(cli_server_version, ["--name", "iknl"])
    assert result.exit_code == 0 and "'server': '3.1.0\\n'" in result.output


@patch(f"{MOD}.ServerContext")
@patch(f"{RT}.containers")
@patch(f"{RT}.ping")
def test_shell_not_running(ping, containers, context):
    containers.list.return_value = []
    context.config_exists.return_value = True
    context.return_value = ctx_mock()
    result = CliRunner().invoke(cli_server_shell, ["--name", "not-running"])
    assert "is not running?" in result.output
