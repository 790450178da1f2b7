"""``vserver`` -- server manager command line interface.

Behavioural spec: reference vantage6/cli/server.py (commands ``start list files new import shell
stop attach version``, the context-injecting decorator ``click_insert_context`` with
``-n -c -e --system/--user``).  The server runtime it launches is a
process managed by vantage6_b200/runtime (not a Docker container) that serves
vantage6_b200/server/app.py.

Deliberate fix: ``vserver version -n NAME`` looks the instance up WITH the ``-server`` suffix
(the reference builds the name without it, so a named lookup never matches: reference
server.py:683; noted as a latent bug in SURVEY.md C28).
"""
from __future__ import annotations

import os
import subprocess
import time
from functools import wraps
from threading import Thread

import click

from .. import runtime as docker
from ..common import check_config_write_permissions, error, info, warning
from ..common import debug as debug_msg
from ..common import prompts as q
from ..common.colors import Fore, Style
from ..common.globals import APPNAME, DEFAULT_DOCKER_REGISTRY, DEFAULT_SERVER_IMAGE, STRING_ENCODING
from ..runtime import LocalRuntime as DockerClient
from ..runtime.addons import (NetworkManager, check_docker_running, get_server_config_name, pull_if_newer,
                              remove_container_if_exists)
from . import __version__
from .configuration_wizard import configuration_wizard, select_configuration_questionaire
from .context import ServerContext, split_db_uri
from .globals import DEFAULT_SERVER_ENVIRONMENT, DEFAULT_SERVER_SYSTEM_FOLDERS
from .rabbitmq.queue_manager import RabbitMQManager
from .utils import check_config_name_allowed

LIST_HEADER = "\nName" + (21 * " ") + "Environments" + (20 * " ") + "Status" + (10 * " ") + "System/User"


def _scope(system_folders: bool) -> str:
    return "system" if system_folders else "user"


def _system_user_options(fn):
    fn = click.option("--user", "system_folders", flag_value=False, default=DEFAULT_SERVER_SYSTEM_FOLDERS)(fn)
    return click.option("--system", "system_folders", flag_value=True)(fn)


def click_insert_context(func):
    """Add ``-n/-c/-e/--system/--user`` and inject a ``ServerContext`` as first argument."""

    @click.option("-n", "--name", default=None, help="name of the configuration you want to use.")
    @click.option("-c", "--config", default=None, help="absolute path to configuration-file; overrides NAME")
    @click.option("-e", "--environment", default=DEFAULT_SERVER_ENVIRONMENT, help="configuration environment to use")
    @_system_user_options
    @wraps(func)
    def func_with_context(name, config, environment, system_folders, *args, **kwargs):
        if config:
            ctx = ServerContext.from_external_config_file(config, environment, system_folders)
        else:
            if not name:
                try:
                    name, environment = select_configuration_questionaire("server", system_folders)
                except Exception:  # noqa: BLE001
                    error("No configurations could be found!")
                    exit(1)
            if not ServerContext.config_exists(name, environment, system_folders):
                scope = _scope(system_folders)
                error(f"Configuration {Fore.RED}{name}{Style.RESET_ALL} with {Fore.RED}{environment}{Style.RESET_ALL} "
                      f"does not exist in the {Fore.RED}{scope}{Style.RESET_ALL} folders!")
                exit(1)
            ServerContext.LOGGING_ENABLED = False
            ctx = ServerContext(name, environment=environment, system_folders=system_folders)
        return func(ctx, *args, **kwargs)

    return func_with_context


@click.group(name="server")
def cli_server():
    """Subcommand `vserver`."""


def print_log_worker(logs_stream):
    for log in logs_stream:
        print(log.decode(STRING_ENCODING), end="")


def _follow_logs(container):
    logs = container.attach(stream=True, logs=True, stdout=True)
    Thread(target=print_log_worker, args=(logs,), daemon=True).start()
    while True:
        try:
            time.sleep(1)
        except KeyboardInterrupt:
            info("Closing log file. Keyboard Interrupt.")
            exit(0)


def _running_server_names(client):
    return [s.name for s in client.containers.list(filters={"label": f"{APPNAME}-type=server"})]


def _database_mounts(ctx, mounts):
    """File-based DB: 'mount' its folder and hand the runtime an absolute sqlite URI through
    ``VANTAGE6_DB_URI`` (reference server.py:177-206)."""
    uri = ctx.config["uri"]
    file_based, db_path = split_db_uri(uri)
    if file_based and db_path:
        if not os.path.isabs(db_path):
            db_path = str(ctx.data_dir / db_path)
        basename, dirname = os.path.basename(db_path), os.path.dirname(db_path)
        os.makedirs(dirname, exist_ok=True)
        mounts.append(docker.types.Mount("/mnt/database/", dirname, type="bind"))
        # environment values are not path-translated by the process runtime (only argv is), so
        # the URI names the host directory that /mnt/database/ stands for
        return {"VANTAGE6_DB_URI": f"sqlite:///{os.path.abspath(dirname)}/{basename}",
                "VANTAGE6_CONFIG_NAME": ctx.config_file_name}
    if not file_based:
        host = uri.split("@")[-1].split("/")[0]
        warning(f"Database could not be transfered, make sure {host} is reachable from the Docker container")
        info("Consider using the docker-compose method to start a server")
    return None


# -------------------------------------------------------------------------------------- start
@cli_server.command(name="start")
@click.option("--ip", default=None, help="ip address to listen on")
@click.option("-p", "--port", default=None, type=int, help="port to listen on")
@click.option("-i", "--image", default=None, help="Server Docker image to use")
@click.option("--rabbitmq-image", default=None, help="RabbitMQ docker image to use")
@click.option("--keep/--auto-remove", default=False, help="Keep image after finishing")
@click.option("--mount-src", default="", help="mount vantage6-master package source")
@click.option("--attach/--detach", default=False, help="Attach server logs to the console after start")
@click_insert_context
def cli_server_start(ctx, ip, port, image, rabbitmq_image, keep, mount_src, attach):
    """Start the server."""
    info("Starting server...")
    info("Finding Docker daemon.")
    docker_client = docker.from_env()
    check_docker_running()
    check_config_name_allowed(ctx.name)

    if f"{APPNAME}-{ctx.name}-{ctx.scope}-server" in _running_server_names(docker_client):
        error(f"Server {Fore.RED}{ctx.name}{Style.RESET_ALL} is already running")
        exit(1)

    if image is None:
        image = ctx.config.get("image", f"{DEFAULT_DOCKER_REGISTRY}/{DEFAULT_SERVER_IMAGE}")
    info(f"Pulling latest server image '{image}'.")
    try:
        pull_if_newer(docker.from_env(), image)
    except Exception:  # noqa: BLE001
        warning("... alas, no dice!")
    else:
        info(" ... success!")

    info("Creating mounts")
    config_file = "/mnt/config.yaml"
    mounts = [docker.types.Mount(config_file, str(ctx.config_file), type="bind")]
    if mount_src:
        mounts.append(docker.types.Mount("/vantage6", os.path.abspath(mount_src), type="bind"))
    environment_vars = _database_mounts(ctx, mounts)

    # a "network" for the server and its sidecars (message queue)
    server_network_mgr = NetworkManager(network_name=f"{APPNAME}-{ctx.name}-{ctx.scope}-network")
    server_network_mgr.create_network(is_internal=False)

    info("Starting RabbitMQ container")
    _start_rabbitmq(ctx, rabbitmq_image, server_network_mgr)

    internal_port = 5000
    cmd = (f"uwsgi --http :{internal_port} --gevent 1000 --http-websockets --master --callable app "
           f"--disable-logging --wsgi-file /vantage6/vantage6-server/vantage6/server/wsgi.py --pyargv {config_file}")
    info(cmd)

    info("Run Docker container")
    port_ = str(port or ctx.config["port"] or 5000)
    container = docker_client.containers.run(
        image, command=cmd, mounts=mounts, detach=True,
        labels={f"{APPNAME}-type": "server", "name": ctx.config_file_name},
        environment=environment_vars, ports={f"{internal_port}/tcp": (ip or ctx.config.get("ip") or "127.0.0.1", port_)},
        name=ctx.docker_container_name, auto_remove=not keep, tty=True, network=server_network_mgr.network_name)
    info(f"Success! container id = {container}")

    if attach:
        _follow_logs(container)


def _start_rabbitmq(ctx: ServerContext, rabbitmq_image: str, network_mgr: NetworkManager) -> None:
    """Start the message-queue sidecar when ``rabbitmq_uri`` is configured."""
    if not ctx.config.get("rabbitmq_uri"):
        warning("Message queue disabled! This means that the server application cannot scale horizontally!")
    else:
        RabbitMQManager(ctx=ctx, network_mgr=network_mgr, image=rabbitmq_image).start()


# --------------------------------------------------------------------------------------- list
@cli_server.command(name="list")
def cli_server_configuration_list():
    """Print the available configurations."""
    client = docker.from_env()
    check_docker_running()
    running_node_names = _running_server_names(client)

    click.echo(LIST_HEADER)
    click.echo("-" * len(LIST_HEADER))
    running = Fore.GREEN + "Online" + Style.RESET_ALL
    stopped = Fore.RED + "Offline" + Style.RESET_ALL
    failed = 0
    for system_folders, tag in ((True, " System "), (False, " User   ")):
        configs, f = ServerContext.available_configurations(system_folders=system_folders)
        failed += len(f)
        for config in configs:
            online = f"{APPNAME}-{config.name}-{_scope(system_folders)}-server" in running_node_names
            status = running if online else stopped
            click.echo(f"{config.name:25}{str(config.available_environments):32}{status:25}{tag}")
    click.echo("-" * 85)
    if failed:
        warning(f"{Fore.RED}Failed imports: {failed}{Style.RESET_ALL}")


# -------------------------------------------------------------------------------------- files
@cli_server.command(name="files")
@click_insert_context
def cli_server_files(ctx):
    """List files locations of a server instance."""
    info(f"Configuration file = {ctx.config_file}")
    info(f"Log file           = {ctx.log_file}")
    info(f"Database           = {ctx.get_database_uri()}")


# ---------------------------------------------------------------------------------------- new
@cli_server.command(name="new")
@click.option("-n", "--name", default=None, help="name of the configutation you want to use.")
@click.option("-e", "--environment", default=DEFAULT_SERVER_ENVIRONMENT, help="configuration environment to use")
@_system_user_options
def cli_server_new(name, environment, system_folders):
    """Create new configuration."""
    if not name:
        name = q.text("Please enter a configuration-name:").ask()
        name_new = name.replace(" ", "-")
        if name != name_new:
            info(f"Replaced spaces from configuration name: {name}")
            name = name_new
    check_config_name_allowed(name)

    try:
        if ServerContext.config_exists(name, environment, system_folders):
            error(f"Configuration {Fore.RED}{name}{Style.RESET_ALL} with environment "
                  f"{Fore.RED}{environment}{Style.RESET_ALL} already exists!")
            exit(1)
    except Exception as e:  # noqa: BLE001
        print(e)
        exit(1)

    if not check_config_write_permissions(system_folders):
        error("Your user does not have write access to all folders. Exiting")
        info(f"Create a new server using '{Fore.GREEN}vserver new --user{Style.RESET_ALL}' instead!")
        exit(1)

    cfg_file = configuration_wizard("server", name, environment=environment, system_folders=system_folders)
    info(f"New configuration created: {Fore.GREEN}{cfg_file}{Style.RESET_ALL}")
    flag = "" if system_folders else "--user"
    info(f"You can start the server by running {Fore.GREEN}vserver start {flag}{Style.RESET_ALL}")


# ------------------------------------------------------------------------------------- import
@cli_server.command(name="import")
@click.argument("file_", type=click.Path(exists=True))
@click.option("--drop-all", is_flag=True, default=False)
@click.option("-i", "--image", default=None, help="Node Docker image to use")
@click.option("--keep/--auto-remove", default=False, help="Keep image after finishing")
@click_insert_context
def cli_server_import(ctx, file_, drop_all, image, keep):
    """Import organizations/collaborations/users and tasks.

    Especially useful for testing purposes.
    """
    info("Starting server...")
    info("Finding Docker daemon.")
    docker_client = docker.from_env()
    check_docker_running()
    check_config_name_allowed(ctx.name)

    if image is None:
        image = ctx.config.get("image", f"{DEFAULT_DOCKER_REGISTRY}/{DEFAULT_SERVER_IMAGE}")
    info(f"Pulling latest server image '{image}'.")
    try:
        docker_client.images.pull(image)
    except Exception:  # noqa: BLE001
        warning("... alas, no dice!")
    else:
        info(" ... success!")

    info("Creating mounts")
    mounts = [docker.types.Mount("/mnt/config.yaml", str(ctx.config_file), type="bind"),
              docker.types.Mount("/mnt/import.yaml", str(os.path.abspath(file_)), type="bind")]
    environment_vars = _database_mounts(ctx, mounts)
    if environment_vars:
        environment_vars.pop("VANTAGE6_CONFIG_NAME", None)

    drop_all_ = "--drop-all" if drop_all else ""
    cmd = f"vserver-local import -c /mnt/config.yaml -e {ctx.environment} {drop_all_} /mnt/import.yaml"
    info(cmd)

    info("Run Docker container")
    container = docker_client.containers.run(
        image, command=cmd, mounts=mounts, detach=True,
        labels={f"{APPNAME}-type": "server", "name": ctx.config_file_name},
        environment=environment_vars, auto_remove=not keep, tty=True)
    logs = container.logs(stream=True, stdout=True)
    Thread(target=print_log_worker, args=(logs,), daemon=False).start()
    info(f"Success! container id = {container.id}")


# -------------------------------------------------------------------------------------- shell
@cli_server.command(name="shell")
@click_insert_context
def cli_server_shell(ctx):
    """Run a iPython shell."""
    docker_client = docker.from_env()
    check_docker_running()
    if ctx.docker_container_name not in _running_server_names(docker_client):
        error(f"Server {Fore.RED}{ctx.name}{Style.RESET_ALL} is not running?")
        return
    try:
        container = docker_client.containers.get(ctx.docker_container_name)
        argv = docker_client._resolve_command("vserver-local shell -c /mnt/config.yaml", container.meta.get("mounts", {}))
        env = dict(os.environ)
        env.update(container.meta.get("environment") or {})
        subprocess.run(argv, env=env)
    except Exception as e:  # noqa: BLE001
        info("Failed to start subprocess...")
        debug_msg(e)


# --------------------------------------------------------------------------------------- stop
@cli_server.command(name="stop")
@click.option("-n", "--name", default=None, help="Configuration name")
@_system_user_options
@click.option("--all", "all_servers", flag_value=True, help="Stop all servers")
def cli_server_stop(name, system_folders, all_servers):
    """Stop a running server"""
    client = docker.from_env()
    check_docker_running()
    running_server_names = _running_server_names(client)
    if not running_server_names:
        warning("No servers are currently running.")
        return

    if all_servers:
        for container_name in running_server_names:
            _stop_server_containers(client, container_name, system_folders)
        return

    if not name:
        container_name = q.select("Select the server you wish to stop:", choices=running_server_names).ask()
    else:
        container_name = f"{APPNAME}-{name}-{_scope(system_folders)}-server"
    if container_name in running_server_names:
        _stop_server_containers(client, container_name, system_folders)
    else:
        error(f"{Fore.RED}{name}{Style.RESET_ALL} is not running!")


def _stop_server_containers(client: DockerClient, container_name: str, system_folders: bool) -> None:
    """Kill the server process and its message-queue sidecar (if any)."""
    client.containers.get(container_name).kill()
    info(f"Stopped the {Fore.GREEN}{container_name}{Style.RESET_ALL} server.")
    config_name = get_server_config_name(container_name, _scope(system_folders))
    rabbit_container_name = f"{APPNAME}-{config_name}-rabbitmq"
    remove_container_if_exists(client, name=rabbit_container_name)
    info(f"Stopped the {Fore.GREEN}{rabbit_container_name}{Style.RESET_ALL} container.")


# ------------------------------------------------------------------------------------- attach
@cli_server.command(name="attach")
@click.option("-n", "--name", default=None, help="configuration name")
@_system_user_options
def cli_server_attach(name, system_folders):
    """Attach the logs from the docker container to the terminal."""
    client = docker.from_env()
    check_docker_running()
    running_server_names = _running_server_names(client)
    if not name:
        name = q.select("Select the server you wish to inspect:", choices=running_server_names).ask()
    else:
        name = f"{APPNAME}-{name}-{_scope(system_folders)}-server"
    if name in running_server_names:
        _follow_logs(client.containers.get(name))
    else:
        error(f"{Fore.RED}{name}{Style.RESET_ALL} was not running!?")


# ------------------------------------------------------------------------------------ version
@cli_server.command(name="version")
@click.option("-n", "--name", default=None, help="configuration name")
@_system_user_options
def cli_server_version(name, system_folders):
    """Returns current version of vantage6 services installed."""
    client = docker.from_env()
    check_docker_running()
    running_server_names = _running_server_names(client)

    if not name:
        if not running_server_names:
            error("No servers are running! You can only check the version for servers that are running")
            exit(1)
        name = q.select("Select the server you wish to inspect:", choices=running_server_names).ask()
    else:
        name = f"{APPNAME}-{name}-{_scope(system_folders)}-server"

    if name in running_server_names:
        version = client.containers.get(name).exec_run(cmd="vserver-local version", stdout=True)
        click.echo({"server": version.output.decode("utf-8"), "cli": __version__})
    else:
        error(f"Server {name} is not running! Cannot provide version...")


if __name__ == "__main__":
    cli_server()
