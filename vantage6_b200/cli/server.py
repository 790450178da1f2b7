"""``vserver`` -- manage central servers on this machine.

Commands: ``start list files new import shell stop attach version``; the context-injecting decorator
:func:`click_insert_context` adds ``-n -c -e --system/--user`` (behavioural spec: reference
vantage6/cli/server.py).  The server that gets launched is a process of the
local runtime (vantage6_b200/runtime) serving vantage6_b200/server/app.py, not a Docker container;
what servers share with nodes lives in ``cli/instance.py``.

Deliberate fix: ``vserver version -n NAME`` looks the instance up WITH the ``-server`` suffix (the
reference builds the name without it, so a named lookup never matches: reference server.py:683;
noted as a latent bug in SURVEY.md C28).
"""
from __future__ import annotations

import os
import subprocess
import sys
import time                                   # noqa: F401  (looked up through this module by instance.follow_logs)
from functools import wraps
from threading import Thread

import click

from .. import runtime as docker
from ..common import check_config_write_permissions, error, info, warning  # noqa: F401
from ..common import debug as debug_msg
from ..common.globals import APPNAME, DEFAULT_DOCKER_REGISTRY, DEFAULT_SERVER_IMAGE
from ..runtime import LocalRuntime as DockerClient
from ..runtime.addons import (NetworkManager, check_docker_running, get_server_config_name, pull_if_newer,  # noqa: F401
                              remove_container_if_exists)
from . import __version__, instance
from .configuration_wizard import configuration_wizard, select_configuration_questionaire
from .context import ServerContext, split_db_uri
from .globals import DEFAULT_SERVER_ENVIRONMENT, DEFAULT_SERVER_SYSTEM_FOLDERS
from .instance import SERVER, green, print_log_worker, red, scope_of  # noqa: F401
from .rabbitmq.queue_manager import RabbitMQManager
from .utils import check_config_name_allowed

_me = sys.modules[__name__]                   # collaborators are resolved on this module at call time
_folders = instance.folders_option(DEFAULT_SERVER_SYSTEM_FOLDERS)
_DEFAULT_IMAGE = f"{DEFAULT_DOCKER_REGISTRY}/{DEFAULT_SERVER_IMAGE}"
_CONFIG_INSIDE = "/mnt/config.yaml"


def click_insert_context(func):
    """Add ``-n/-c/-e/--system/--user`` to a command and hand it a ``ServerContext`` as first argument."""

    @instance.name_option("name of the configuration you want to use.")
    @instance.config_option()
    @instance.environment_option(DEFAULT_SERVER_ENVIRONMENT)
    @_folders
    @wraps(func)
    def with_context(name, config, environment, system_folders, *args, **kwargs):
        if config:
            return func(ServerContext.from_external_config_file(config, environment, system_folders), *args, **kwargs)
        if not name:
            try:
                name, environment = select_configuration_questionaire("server", system_folders)
            except Exception:  # noqa: BLE001
                error("No configurations could be found!")
                exit(1)
        if not ServerContext.config_exists(name, environment, system_folders):
            error(f"Configuration {red(name)} with {red(environment)} "
                  f"does not exist in the {red(scope_of(system_folders))} folders!")
            exit(1)
        ServerContext.LOGGING_ENABLED = False
        return func(ServerContext(name, environment=environment, system_folders=system_folders), *args, **kwargs)

    return with_context


@click.group(name="server")
def cli_server():
    """Subcommand `vserver`."""


def _running_server_names(client):
    return SERVER.running(client)


def _refresh_image(ctx, requested, puller):
    image = instance.choose_image(_me, ctx, requested, _DEFAULT_IMAGE, "server", tail=".")
    try:
        puller(image)
    except Exception:  # noqa: BLE001 -- best effort, a local image may do
        instance.report_pull(False, lead="")
    else:
        instance.report_pull(True)
    return image


def _database_mounts(ctx, mounts):
    """A file-based database is made available as ``/mnt/database/`` and the runtime gets an absolute sqlite
    URI through ``VANTAGE6_DB_URI``; for any other database only a reachability warning can be given
    (reference server.py:177-206).  Returns the environment for the launched process (or None)."""
    uri = ctx.config["uri"]
    is_file, db_file = split_db_uri(uri)
    if is_file and db_file:
        db_file = db_file if os.path.isabs(db_file) else str(ctx.data_dir / db_file)
        folder = os.path.dirname(db_file)
        os.makedirs(folder, exist_ok=True)
        mounts.append(docker.types.Mount("/mnt/database/", folder, type="bind"))
        # environment values are not path-translated by the process runtime (only argv is): the URI names
        # the host directory that /mnt/database/ stands for
        return {"VANTAGE6_DB_URI": f"sqlite:///{os.path.abspath(folder)}/{os.path.basename(db_file)}",
                "VANTAGE6_CONFIG_NAME": ctx.config_file_name}
    if not is_file:
        host = uri.split("@")[-1].split("/")[0]
        warning(f"Database could not be transfered, make sure {host} is reachable from the Docker container")
        info("Consider using the docker-compose method to start a server")
    return None


# ------------------------------------------------------------------------------------------ start
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
    rt = docker.from_env()
    check_docker_running()
    check_config_name_allowed(ctx.name)
    if f"{APPNAME}-{ctx.name}-{ctx.scope}-server" in _running_server_names(rt):
        error(f"Server {red(ctx.name)} is already running")
        exit(1)

    image = _refresh_image(ctx, image, lambda img: pull_if_newer(docker.from_env(), img))

    info("Creating mounts")
    mounts = [docker.types.Mount(_CONFIG_INSIDE, str(ctx.config_file), type="bind")]
    if mount_src:
        mounts.append(docker.types.Mount("/vantage6", os.path.abspath(mount_src), type="bind"))
    process_env = _database_mounts(ctx, mounts)

    # one "network" for the server and its sidecars (the message queue)
    network = NetworkManager(network_name=f"{APPNAME}-{ctx.name}-{ctx.scope}-network")
    network.create_network(is_internal=False)
    info("Starting RabbitMQ container")
    _start_rabbitmq(ctx, rabbitmq_image, network)

    inside_port = 5000
    cmd = (f"uwsgi --http :{inside_port} --gevent 1000 --http-websockets --master --callable app "
           f"--disable-logging --wsgi-file /vantage6/vantage6-server/vantage6/server/wsgi.py --pyargv {_CONFIG_INSIDE}")
    info(cmd)

    info("Run Docker container")
    listen = (ip or ctx.config.get("ip") or "127.0.0.1", str(port or ctx.config["port"] or 5000))
    container = rt.containers.run(
        image, command=cmd, mounts=mounts, detach=True, environment=process_env,
        labels={f"{APPNAME}-type": "server", "name": ctx.config_file_name}, ports={f"{inside_port}/tcp": listen},
        name=ctx.docker_container_name, auto_remove=not keep, tty=True, network=network.network_name)
    info(f"Success! container id = {container}")
    if attach:
        instance.follow_logs(_me, container, stdout=True)


def _start_rabbitmq(ctx: ServerContext, rabbitmq_image: str, network_mgr: NetworkManager) -> None:
    """The message-queue sidecar runs only when ``rabbitmq_uri`` is configured."""
    if ctx.config.get("rabbitmq_uri"):
        RabbitMQManager(ctx=ctx, network_mgr=network_mgr, image=rabbitmq_image).start()
    else:
        warning("Message queue disabled! This means that the server application cannot scale horizontally!")


# ------------------------------------------------------------------------------ list / files / new
@cli_server.command(name="list")
def cli_server_configuration_list():
    """Print the available configurations."""
    instance.show_table(_me, SERVER)


@cli_server.command(name="files")
@click_insert_context
def cli_server_files(ctx):
    """List files locations of a server instance."""
    for title, value in (("Configuration file", ctx.config_file), ("Log file          ", ctx.log_file),
                         ("Database          ", ctx.get_database_uri())):
        info(f"{title} = {value}")


@cli_server.command(name="new")
@instance.name_option("name of the configutation you want to use.")
@instance.environment_option(DEFAULT_SERVER_ENVIRONMENT)
@_folders
def cli_server_new(name, environment, system_folders):
    """Create new configuration."""
    name = instance.ask_configuration_name(_me, name, always_normalise=False)
    check_config_name_allowed(name)
    try:
        taken = ServerContext.config_exists(name, environment, system_folders)
    except Exception as e:  # noqa: BLE001
        print(e)
        exit(1)
    if taken:
        error(f"Configuration {red(name)} with environment {red(environment)} already exists!")
        exit(1)
    instance.require_write_access(_me, system_folders,
                                  hint=f"Create a new server using '{green('vserver new --user')}' instead!")
    cfg_file = configuration_wizard("server", name, environment=environment, system_folders=system_folders)
    instance.announce_new_configuration(SERVER, cfg_file, "" if system_folders else "--user")


# ----------------------------------------------------------------------------------------- import
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
    rt = docker.from_env()
    check_docker_running()
    check_config_name_allowed(ctx.name)
    image = _refresh_image(ctx, image, rt.images.pull)

    info("Creating mounts")
    mounts = [docker.types.Mount(_CONFIG_INSIDE, str(ctx.config_file), type="bind"),
              docker.types.Mount("/mnt/import.yaml", str(os.path.abspath(file_)), type="bind")]
    process_env = _database_mounts(ctx, mounts)
    if process_env:
        process_env.pop("VANTAGE6_CONFIG_NAME", None)

    cmd = (f"vserver-local import -c {_CONFIG_INSIDE} -e {ctx.environment} "
           f"{'--drop-all' if drop_all else ''} /mnt/import.yaml")
    info(cmd)

    info("Run Docker container")
    container = rt.containers.run(image, command=cmd, mounts=mounts, detach=True, environment=process_env,
                                  labels={f"{APPNAME}-type": "server", "name": ctx.config_file_name},
                                  auto_remove=not keep, tty=True)
    Thread(target=print_log_worker, args=(container.logs(stream=True, stdout=True),), daemon=False).start()
    info(f"Success! container id = {container.id}")


# ------------------------------------------------------------------------------------------ shell
@cli_server.command(name="shell")
@click_insert_context
def cli_server_shell(ctx):
    """Run a iPython shell."""
    rt = docker.from_env()
    check_docker_running()
    if ctx.docker_container_name not in _running_server_names(rt):
        error(f"Server {red(ctx.name)} is not running?")
        return
    try:
        running = rt.containers.get(ctx.docker_container_name)
        argv = rt._resolve_command(f"vserver-local shell -c {_CONFIG_INSIDE}", running.meta.get("mounts", {}))
        subprocess.run(argv, env={**os.environ, **(running.meta.get("environment") or {})})
    except Exception as e:  # noqa: BLE001
        info("Failed to start subprocess...")
        debug_msg(e)


# ---------------------------------------------------------------------- stop / attach / version
def _stop_server_containers(client: DockerClient, container_name: str, system_folders: bool) -> None:
    """End the server process and its message-queue sidecar (if any)."""
    client.containers.get(container_name).kill()
    info(f"Stopped the {green(container_name)} server.")
    sidecar = f"{APPNAME}-{get_server_config_name(container_name, scope_of(system_folders))}-rabbitmq"
    remove_container_if_exists(client, name=sidecar)
    info(f"Stopped the {green(sidecar)} container.")


@cli_server.command(name="stop")
@instance.name_option("Configuration name")
@_folders
@click.option("--all", "all_servers", flag_value=True, help="Stop all servers")
def cli_server_stop(name, system_folders, all_servers):
    """Stop a running server"""
    instance.stop(_me, SERVER, name, system_folders, bool(all_servers),
                  halt=lambda rt, runtime_name: _stop_server_containers(rt, runtime_name, system_folders))


@cli_server.command(name="attach")
@instance.name_option()
@_folders
def cli_server_attach(name, system_folders):
    """Attach the logs from the docker container to the terminal."""
    instance.attach(_me, SERVER, name, system_folders, stdout=True)


@cli_server.command(name="version")
@instance.name_option()
@_folders
def cli_server_version(name, system_folders):
    """Returns current version of vantage6 services installed."""
    instance.version(_me, SERVER, name, system_folders, __version__)


if __name__ == "__main__":
    cli_server()
