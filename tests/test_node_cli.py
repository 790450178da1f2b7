"""vnode CLI unit tests: CliRunner + mocks for the process runtime, the context and the prompts
(the reference's single test tier; cases mirror reference
tests/test_node_cli.py and add the ones it leaves untested: remove, version)."""
import contextlib
from io import StringIO
from pathlib import Path
from unittest.mock import MagicMock, patch

import pytest
from click.testing import CliRunner

from vantage6_b200.cli.globals import APPNAME
from vantage6_b200.cli.node import (cli_node_attach, cli_node_clean, cli_node_create_private_key, cli_node_files,
                                    cli_node_list, cli_node_new_configuration, cli_node_remove, cli_node_start,
                                    cli_node_stop, cli_node_version, create_client_and_authenticate, print_log_worker)
from vantage6_b200.common import STRING_ENCODING
from vantage6_b200.runtime import APIError

RT = "vantage6_b200.runtime.LocalRuntime"
MOD = "vantage6_b200.cli.node"


def container(name):
    c = MagicMock()
    c.name = name
    return c


@patch(f"{RT}.ping")
def test_list_runtime_not_running(ping):
    """An error + exit 1 when the runtime is unusable (reference: docker not running)."""
    ping.side_effect = Exception("Boom!")
    result = CliRunner().invoke(cli_node_list, [])
    assert result.exit_code == 1


@patch(f"{MOD}.NodeContext.available_configurations")
@patch(f"{RT}.ping")
@patch(f"{RT}.containers")
def test_list_golden_output(containers, ping, available_configurations):
    """Byte-exact table (reference tests/test_node_cli.py:80-87)."""
    ping.return_value = True
    containers.list.return_value = [container(f"{APPNAME}-iknl-user")]

    def side_effect(system_folders):
        config = MagicMock(available_environments=["Application"])
        config.name = "iknl"
        return [[config], []]

    available_configurations.side_effect = side_effect
    result = CliRunner().invoke(cli_node_list, [])
    assert result.exit_code == 0
    assert result.output == (
        "\nName                     Environments                    Status          System/User\n"
        "-------------------------------------------------------------------------------------\n"
        "iknl                     ['Application']                 Offline          System \n"
        "iknl                     ['Application']                 Online           User   \n"
        "-------------------------------------------------------------------------------------\n")


@patch(f"{MOD}.configuration_wizard")
@patch(f"{MOD}.check_config_write_permissions")
@patch(f"{MOD}.NodeContext")
def test_new_config(context, permissions, wizard):
    context.config_exists.return_value = False
    permissions.return_value = True
    wizard.return_value = "/some/file/path"
    result = CliRunner().invoke(cli_node_new_configuration, ["--name", "some-name", "--environment", "application"])
    assert result.output[:6] == "[info]"
    assert result.exit_code == 0


@patch(f"{MOD}.configuration_wizard")
@patch(f"{MOD}.check_config_write_permissions", return_value=True)
@patch(f"{MOD}.NodeContext")
def test_new_config_replace_whitespace_in_name(context, _perm, _wiz):
    context.config_exists.return_value = False
    result = CliRunner().invoke(cli_node_new_configuration, ["--name", "some name", "--environment", "application"])
    assert result.output[:60] == "[info]  - Replaced spaces from configuration name: some-name"


def test_new_config_invalid_name():
    result = CliRunner().invoke(cli_node_new_configuration, ["--name", "bad$name", "--environment", "application"])
    assert result.output.startswith("[error]")
    assert result.exit_code == 1


@patch(f"{MOD}.NodeContext")
def test_new_config_already_exists(context):
    context.config_exists.return_value = True
    result = CliRunner().invoke(cli_node_new_configuration, ["--name", "some-name", "--environment", "application"])
    assert result.output[:7] == "[error]"
    assert result.exit_code == 1


@patch(f"{MOD}.check_config_write_permissions")
@patch(f"{MOD}.NodeContext")
def test_new_write_permissions(context, permissions):
    context.config_exists.return_value = False
    permissions.return_value = False
    result = CliRunner().invoke(cli_node_new_configuration, ["--name", "some-name", "--environment", "application"])
    assert result.output[:7] == "[error]"
    assert result.exit_code == 1


@patch(f"{MOD}.NodeContext")
@patch(f"{MOD}.select_configuration_questionaire")
def test_files(select_config, context):
    context.config_exists.return_value = True
    context.return_value = MagicMock(config_file="/file.yaml", log_file="/log.log", data_dir="/dir")
    context.return_value.databases.items.return_value = [["label", "/file.db"]]
    select_config.return_value = ["iknl", "application"]
    result = CliRunner().invoke(cli_node_files, [])
    assert result.output[:6] == "[info]"
    assert "label" in result.output
    assert result.exit_code == 0


@patch(f"{MOD}.NodeContext")
def test_files_non_existing_config(context):
    context.config_exists.return_value = False
    result = CliRunner().invoke(cli_node_files, ["--name", "non-existing"])
    assert result.output[:7] == "[error]"
    assert result.exit_code != 0


def _start_ctx():
    ctx = MagicMock(data_dir=Path("data"), log_dir=Path("logs"), config_dir=Path("configs"),
                    databases={"default": "data.csv"})
    ctx.get_data_file.return_value = "data.csv"
    ctx.name = "some-name"
    ctx.config = {"encryption": {}}
    ctx.docker_container_name = f"{APPNAME}-some-name-user"
    ctx.docker_volume_name = "vol"
    ctx.docker_vpn_volume_name = "vpn-vol"
    return ctx


@patch(f"{RT}.volumes")
@patch(f"{MOD}.pull_if_newer")
@patch(f"{MOD}.NodeContext")
@patch(f"{RT}.containers")
@patch(f"{RT}.ping")
def test_start(ping, containers, context, pull, volumes):
    containers.list.return_value = []
    volume = MagicMock()
    volume.name = "data-vol-name"
    volumes.create.return_value = volume
    context.config_exists.return_value = True
    context.return_value = _start_ctx()
    runner = CliRunner()
    with runner.isolated_filesystem():
        result = runner.invoke(cli_node_start, ["--name", "some-name", "--gpu", "3"])
    assert result.exit_code == 0, result.output
    kwargs = containers.run.call_args.kwargs
    assert kwargs["command"].startswith("vnode-local start -c /mnt/config/some-name.yaml -n some-name -e application")
    assert kwargs["labels"][f"{APPNAME}-type"] == "node"
    assert kwargs["environment"]["V6_GPU"] == "3"
    assert kwargs["environment"]["DATA_VOLUME_NAME"] == "data-vol-name"
    assert "DEFAULT_DATABASE_URI" in kwargs["environment"]


@patch(f"{RT}.volumes")
@patch(f"{MOD}.NodeContext")
@patch(f"{RT}.containers")
@patch(f"{RT}.ping")
def test_start_already_running(ping, containers, context, volumes):
    containers.list.return_value = [container(f"{APPNAME}-some-name-user")]
    context.config_exists.return_value = True
    context.return_value = _start_ctx()
    result = CliRunner().invoke(cli_node_start, ["--name", "some-name"])
    assert "already running" in result.output
    assert result.exit_code == 1


@patch(f"{RT}.containers")
@patch(f"{RT}.ping")
def test_stop(ping, containers):
    containers.list.return_value = [container(f"{APPNAME}-iknl-user")]
    result = CliRunner().invoke(cli_node_stop, ["--name", "iknl"])
    assert result.output == "[info]  - Stopped the vantage6-iknl-user Node.\n"
    assert result.exit_code == 0
    containers.get.return_value.stop.assert_called_once()


@patch(f"{RT}.containers")
@patch(f"{RT}.ping")
def test_stop_all_kills(ping, containers):
    containers.list.return_value = [container(f"{APPNAME}-a-user"), container(f"{APPNAME}-b-user")]
    result = CliRunner().invoke(cli_node_stop, ["--all"])
    assert result.exit_code == 0
    assert containers.get.return_value.kill.call_count == 2


@patch(f"{RT}.containers")
@patch(f"{RT}.ping")
def test_stop_nothing_running(ping, containers):
    containers.list.return_value = []
    result = CliRunner().invoke(cli_node_stop, ["--name", "iknl"])
    assert result.output.startswith("[warn]")


@patch(f"{MOD}.time")
@patch(f"{MOD}.print_log_worker")
@patch(f"{RT}.containers")
@patch(f"{RT}.ping")
def test_attach(ping, containers, log_worker, time_):
    containers.list.return_value = [container(f"{APPNAME}-iknl-user")]
    log_worker.return_value = ""
    time_.sleep.side_effect = KeyboardInterrupt()
    result = CliRunner().invoke(cli_node_attach, ["--name", "iknl"])
    assert result.output == "[info]  - Closing log file. Keyboard Interrupt.\n"
    assert result.exit_code == 0


@patch(f"{MOD}.q")
@patch(f"{RT}.volumes")
@patch(f"{RT}.ping")
def test_clean(ping, volumes, q):
    volume1 = MagicMock()
    volume1.name = "some-name-tmpvol"
    other = MagicMock()
    other.name = "keep-me"
    volumes.list.return_value = [volume1, other]
    q.confirm.return_value.ask.return_value = True
    result = CliRunner().invoke(cli_node_clean)
    assert result.exit_code == 0
    volume1.remove.assert_called_once()
    other.remove.assert_not_called()


@patch(f"{MOD}.q")
@patch(f"{RT}.volumes")
@patch(f"{RT}.ping")
def test_clean_api_error_exits_1(ping, volumes, q):
    volume1 = MagicMock()
    volume1.name = "some-name-tmpvol"
    volume1.remove.side_effect = APIError("in use")
    volumes.list.return_value = [volume1]
    q.confirm.return_value.ask.return_value = True
    result = CliRunner().invoke(cli_node_clean)
    assert result.exit_code == 1


@patch(f"{MOD}.create_client_and_authenticate")
@patch(f"{MOD}.NodeContext")
def test_create_private_key(context, client):
    context.config_exists.return_value = True
    runner = CliRunner()
    with runner.isolated_filesystem():
        context.return_value.type_data_folder.return_value = Path(".")
        context.return_value.config = {"encryption": {}}
        client.return_value = MagicMock(whoami=MagicMock(organization_name="Test", organization_id=7))
        result = runner.invoke(cli_node_create_private_key, ["--name", "application"])
        assert Path("privkey_Test.pem").exists()
    assert result.exit_code == 0, result.output
    args, kwargs = client.return_value.request.call_args
    assert args[0] == "/organization/7" and kwargs["method"] == "patch" and "public_key" in kwargs["json"]


@patch(f"{MOD}.RSACryptor")
@patch(f"{MOD}.create_client_and_authenticate")
@patch(f"{MOD}.NodeContext")
def test_create_private_key_overwite(context, client, cryptor):
    context.config_exists.return_value = True
    cryptor.create_public_key_bytes.return_value = b""
    runner = CliRunner()
    with runner.isolated_filesystem():
        context.return_value.type_data_folder.return_value = Path(".")
        context.return_value.config = {"encryption": {}}
        client.return_value = MagicMock(whoami=MagicMock(organization_name="Test"))
        Path("privkey_iknl.pem").write_text("does-not-matter")
        result = runner.invoke(cli_node_create_private_key,
                               ["--name", "application", "--overwrite", "--organization-name", "iknl"])
        assert Path("privkey_iknl.pem").exists()
    assert result.exit_code == 0
    cryptor.create_new_rsa_key.assert_called_once()


@patch(f"{MOD}.RSACryptor")
@patch(f"{MOD}.NodeContext")
def test_create_private_key_keeps_existing_without_overwrite(context, cryptor):
    context.config_exists.return_value = True
    cryptor.create_public_key_bytes.return_value = b""
    runner = CliRunner()
    with runner.isolated_filesystem():
        context.return_value.type_data_folder.return_value = Path(".")
        context.return_value.config = {"encryption": {}}
        Path("privkey_iknl.pem").write_text("does-not-matter")
        result = runner.invoke(cli_node_create_private_key,
                               ["--name", "application", "--organization-name", "iknl", "--no-upload"])
    assert "Continuing with existing key instead!" in result.output
    cryptor.create_new_rsa_key.assert_not_called()
    assert result.exit_code == 0


@patch(f"{MOD}.NodeContext")
def test_create_private_key_config_not_found(context):
    context.config_exists.return_value = False
    result = CliRunner().invoke(cli_node_create_private_key, ["--name", "application"])
    assert result.exit_code == 1


@patch(f"{MOD}.q")
@patch(f"{MOD}.NodeContext")
@patch(f"{RT}.volumes")
@patch(f"{RT}.containers")
@patch(f"{RT}.ping")
def test_remove(ping, containers, volumes, context, q):
    containers.list.return_value = []
    context.config_exists.return_value = True
    q.confirm.return_value.ask.return_value = True
    runner = CliRunner()
    with runner.isolated_filesystem():
        Path("cfg.yaml").write_text("x")
        Path("node.log").write_text("x")
        ctx = MagicMock(config_file="cfg.yaml", log_file="node.log", data_dir=".", docker_volume_name="vantage6-n-user-vol",
                        docker_vpn_volume_name="vantage6-n-user-vpn-vol")
        ctx.log.handlers, ctx.log.root.handlers = [], []
        context.return_value = ctx
        v1, v2 = MagicMock(), MagicMock()
        v1.name, v2.name = "vantage6-n-user-vol", "unrelated"
        volumes.list.return_value = [v1, v2]
        result = runner.invoke(cli_node_remove, ["--name", "n"])
        assert not Path("cfg.yaml").exists() and not Path("node.log").exists()
    assert result.exit_code == 0, result.output
    v1.remove.assert_called_once()
    v2.remove.assert_not_called()


@patch(f"{MOD}.NodeContext")
@patch(f"{RT}.containers")
@patch(f"{RT}.ping")
def test_remove_refuses_running_node(ping, containers, context):
    context.config_exists.return_value = True
    containers.list.return_value = [container(f"{APPNAME}-n-user")]
    result = CliRunner().invoke(cli_node_remove, ["--name", "n"])
    assert "still running" in result.output
    assert result.exit_code == 1


@patch(f"{RT}.containers")
@patch(f"{RT}.ping")
def test_version(ping, containers):
    containers.list.return_value = [container(f"{APPNAME}-iknl-user")]
    containers.get.return_value.exec_run.return_value = MagicMock(output=b"3.1.0\n")
    result = CliRunner().invoke(cli_node_version, ["--name", "iknl"])
    assert "'node': '3.1.0\\n'" in result.output and "'cli'" in result.output
    assert result.exit_code == 0


@patch(f"{MOD}.q")
@patch(f"{MOD}.Client")
def test_client(client, q):
    ctx = MagicMock(config={"server_url": "localhost", "port": 5000, "api_path": ""})
    cli = create_client_and_authenticate(ctx)
    cli.authenticate.assert_called_once()


@patch(f"{MOD}.q")
@patch(f"{MOD}.Client")
def test_client_authentication_error_exits(client, q):
    client.return_value.authenticate.side_effect = Exception("bad credentials")
    ctx = MagicMock(config={"server_url": "localhost", "port": 5000, "api_path": ""})
    with pytest.raises(SystemExit):
        with contextlib.redirect_stdout(StringIO()):
            create_client_and_authenticate(ctx)


def test_print_log_worker():
    stream = [b"hello\n", "wörld\n".encode(STRING_ENCODING)]
    buf = StringIO()
    with contextlib.redirect_stdout(buf):
        print_log_worker(iter(stream))
    assert buf.getvalue() == "hello\nwörld\n"
