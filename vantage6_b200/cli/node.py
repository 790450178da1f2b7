"""``vnode`` -- manage federated nodes on this machine.

Commands: ``list new files start stop attach create-private-key clean remove version`` with the
options, defaults, console messages and exit codes of the reference CLI (behavioural spec:
reference vantage6/cli/node.py).  Two things are different by design:

* what gets launched -- a node is ONE process pinned to ONE B200 (``--gpu K`` -> ``V6_GPU=K``, all
  GPUs stay visible so peers can be mapped over NVLink) managed by the process runtime
  (vantage6_b200/runtime) instead of a Docker container;
* how the file is organised -- everything nodes share with servers lives in ``cli/instance.py``;
  ``start`` assembles a :class:`~vantage6_b200.cli.instance.LaunchPlan` in small steps.
"""
from __future__ import annotations

import itertools
import os.path
import sys
import time                                   # noqa: F401  (looked up through this module by instance.follow_logs)
from pathlib import Path

import click

from .. import runtime as docker              # docker-SDK-shaped process runtime
from ..client import Client
from ..client.encryption import RSACryptor
from ..common import bytes_to_base64s, check_config_write_permissions, debug, error, info, warning  # noqa: F401
from ..common import prompts as q
from ..common.colors import Fore, Style
from ..common.globals import APPNAME, DEFAULT_DOCKER_REGISTRY, DEFAULT_NODE_IMAGE, VPN_CONFIG_FILE
from ..runtime.addons import check_docker_running, pull_if_newer, remove_container_if_exists  # noqa: F401
from . import __version__, instance
from .configuration_wizard import configuration_wizard, select_configuration_questionaire
from .context import NodeContext
from .globals import DEFAULT_NODE_ENVIRONMENT as N_ENV
from .globals import DEFAULT_NODE_SYSTEM_FOLDERS as N_FOL
from .instance import NODE, LaunchPlan, green, print_log_worker, red  # noqa: F401
from .utils import check_config_name_allowed, check_if_docker_deamon_is_running

_me = sys.modules[__name__]                   # collaborators are resolved on this module at call time
_folders = instance.folders_option(N_FOL)
_environment = instance.environment_option(N_ENV)


@click.group(name="node")
def cli_node():
    """Subcommand `vnode`."""


def find_running_node_names(client):
    return NODE.running(client)


# ------------------------------------------------------------------------------ list / new / files
@cli_node.command(name="list")
def cli_node_list():
    """Lists all nodes in the default configuration directory."""
    instance.show_table(_me, NODE)


@cli_node.command(name="new")
@instance.name_option(None)
@click.option("-e", "--environment", default="", help="configuration environment to use")
@_folders
def cli_node_new_configuration(name, environment, system_folders):
    """Create a new configuration file.

    Checks if the configuration already exists. If this is not the case a questionaire is
    invoked to create a new configuration file.
    """
    name = instance.ask_configuration_name(_me, name, always_normalise=True)
    check_config_name_allowed(name)
    environment = environment or q.select("Please select the environment you want to configure:",
                                          list(instance.ENVIRONMENTS)).ask()
    if NodeContext.config_exists(name, environment, system_folders):
        error(f"Configuration {name} and environment{environment} already exists!")
        exit(1)
    instance.require_write_access(_me, system_folders)
    cfg_file = configuration_wizard("node", name, environment, system_folders)
    instance.announce_new_configuration(NODE, cfg_file, "--system" if system_folders else "")


@cli_node.command(name="files")
@instance.name_option()
@_environment
@_folders
def cli_node_files(name, environment, system_folders):
    """Prints location important files.

    If the specified configuration cannot be found, it exits. Otherwise it returns the
    absolute path to the output.
    """
    name, environment = select_node(name, environment, system_folders)
    ctx = NodeContext(name, environment=environment, system_folders=system_folders)
    for title, value in (("Configuration file", ctx.config_file), ("Log file          ", ctx.log_file),
                         ("data folders      ", ctx.data_dir)):
        info(f"{title} = {value}")
    info("Database labels and files")
    for label, path in ctx.databases.items():
        info(f" - {label:15} = {path}")


# ------------------------------------------------------------------------------------------ start
def _context_for_start(name, config, environment, system_folders):
    """``-c FILE`` wins; otherwise pick / create the named configuration (wizard fallback)."""
    NodeContext.LOGGING_ENABLED = False
    if config:
        return Path(config).stem, environment, NodeContext(Path(config).stem, environment, system_folders, config)
    if not name:
        name, environment = select_configuration_questionaire("node", system_folders)
    if not NodeContext.config_exists(name, environment, system_folders):
        wanted = q.confirm(f"Configuration '{name}' using environment '{environment}' does not exist.\n"
                           "  Do you want to create this config now?").ask()
        if not wanted:
            error("Config file couldn't be loaded")
            sys.exit(0)
        configuration_wizard("node", name, environment, system_folders)
    return name, environment, NodeContext(name, environment, system_folders)


def _refuse_second_instance(rt, name, system_folders) -> None:
    if NODE.runtime_name(name, system_folders) in [n.name for n in rt.containers.list(filters={"label": NODE.label})]:
        error(f"Node {red(name)} is already running")
        exit(1)


def _plan_private_key(plan: LaunchPlan, ctx) -> None:
    """config -> default file name -> ``PRIVATE_KEY`` environment override (reference node.py:331-348)."""
    configured = ctx.config.get("encryption", {}).get("private_key") or "private_key.pem"
    key_path = Path(ctx.get_data_file(os.environ.get("PRIVATE_KEY", configured)))
    if key_path.exists():
        plan.bind("/mnt/private_key.pem", key_path)
        plan.environment["PRIVATE_KEY"] = str(key_path)
    else:
        warning(f"private key file provided {key_path}, but does not exists")


def _plan_databases(plan: LaunchPlan, ctx, force_db_mount: bool) -> None:
    """Every labelled database becomes ``{LABEL}_DATABASE_URI``; files are also bound at /mnt/{label}.csv."""
    info("Setting up databases")
    for label, uri in ctx.databases.items():
        info(f"  Processing database '{label}:{uri}'")
        key = f"{label.upper()}_DATABASE_URI"
        if Path(uri).exists() or force_db_mount:
            debug("  - file-based database added")
            plan.environment[key] = str(Path(uri).resolve())
            plan.bind(f"/mnt/{label}.csv", uri)
        else:
            debug("  - non file-based database added")
            plan.environment[key] = uri
        if label == "default":
            plan.environment["DATABASE_URI"] = plan.environment[key]        # vantage6 2.x name


@cli_node.command(name="start")
@instance.name_option()
@instance.config_option()
@_environment
@_folders
@click.option("-i", "--image", default=None, help="Node Docker image to use")
@click.option("--keep/--auto-remove", default=False, help="Keep image after finishing")
@click.option("--force-db-mount", is_flag=True,
              help="Skip the check of the existence of the DB (always try to mount)")
@click.option("--attach/--detach", default=False, help="Attach node logs to the console after start")
@click.option("--mount-src", default="", help="mount vantage6-master package source")
@click.option("--gpu", default=None, type=int, help="B200 device index this node is pinned to")
def cli_node_start(name, config, environment, system_folders, image, keep, mount_src, attach, force_db_mount,
                   gpu=None):
    """Start the node instance.

    If no name or config is specified the default.yaml configuation is used. In case the
    configuration file not excists, a questionaire is invoked to create one.
    """
    info("Starting node...")
    info("Finding Docker deamon")
    rt = docker.from_env()
    check_docker_running()

    name, environment, ctx = _context_for_start(name, config, environment, system_folders)
    check_config_name_allowed(ctx.name)
    _refuse_second_instance(rt, name, system_folders)

    info("Checking that data and log dirs exist")
    for folder in (ctx.data_dir, ctx.log_dir):
        folder.mkdir(parents=True, exist_ok=True)

    plan = LaunchPlan(name=ctx.docker_container_name, keep=keep)
    plan.image = instance.choose_image(_me, ctx, image, f"{DEFAULT_DOCKER_REGISTRY}/{DEFAULT_NODE_IMAGE}", "node")
    try:
        pull_if_newer(docker.from_env(), plan.image)
    except Exception:  # noqa: BLE001 -- best effort, a local image may do
        instance.report_pull(False)
    else:
        instance.report_pull(True)

    info("Creating Docker data volume")
    data_volume = rt.volumes.create(ctx.docker_volume_name)
    vpn_volume = rt.volumes.create(ctx.docker_vpn_volume_name)
    plan.environment.update(DATA_VOLUME_NAME=data_volume.name, VPN_VOLUME_NAME=vpn_volume.name)

    info("Creating file & folder mounts")
    plan.bind("/mnt/log", ctx.log_dir)
    plan.bind("/mnt/data", data_volume.name)
    plan.bind("/mnt/vpn", vpn_volume.name)
    plan.bind("/mnt/config", ctx.config_dir)
    if mount_src:
        plan.bind("/vantage6", os.path.abspath(mount_src))
    _plan_private_key(plan, ctx)
    _plan_databases(plan, ctx, force_db_mount)

    if gpu is None and ctx.config.get("gpu") is not None:
        gpu = int(ctx.config.get("gpu"))
    if gpu is not None:
        plan.environment["V6_GPU"] = str(gpu)
        info(f"Pinning node to GPU {gpu}")

    plan.command = (f"vnode-local start -c /mnt/config/{name}.yaml -n {name} -e {environment} "
                    f"--dockerized {'--system' if system_folders else '--user'}")
    plan.labels = {f"{APPNAME}-type": "node", "system": str(system_folders), "name": ctx.config_file_name}

    info("Running Docker container")
    remove_container_if_exists(docker_client=rt, name=plan.name)
    container = rt.containers.run(plan.image, command=plan.command, volumes=plan.volume_specs(), detach=True,
                                  labels=plan.labels, environment=plan.environment, name=plan.name,
                                  auto_remove=not plan.keep, tty=True)
    info(f"Success! container id = {container}")
    if attach:
        instance.follow_logs(_me, container)


# ---------------------------------------------------------------------- stop / attach / version
def _halt_node(rt, runtime_name: str, graceful: bool) -> None:
    target = rt.containers.get(runtime_name)
    target.stop() if graceful else target.kill()          # stop(): 10 s to exit by itself, then killed
    info(f"Stopped the {green(runtime_name)} Node.")


@cli_node.command(name="stop")
@instance.name_option()
@_folders
@click.option("--all", "all_nodes", flag_value=True)
def cli_node_stop(name, system_folders, all_nodes):
    """Stop a running container."""
    instance.stop(_me, NODE, name, system_folders, bool(all_nodes),
                  halt=lambda rt, runtime_name: _halt_node(rt, runtime_name, graceful=not all_nodes))


@cli_node.command(name="attach")
@instance.name_option()
@_folders
def cli_node_attach(name, system_folders):
    """Attach the logs from the docker container to the terminal."""
    instance.attach(_me, NODE, name, system_folders)


@cli_node.command(name="version")
@instance.name_option()
@_folders
def cli_node_version(name, system_folders):
    """Returns current version of vantage6 services installed."""
    instance.version(_me, NODE, name, system_folders, __version__)


# ----------------------------------------------------------------------------- create-private-key
def _existing_or_new_key(key_file: Path, overwrite: bool):
    """The RSA private key for ``key_file``: kept when present (unless ``--overwrite``), else generated."""
    present = key_file.exists()
    if present:
        warning(f"File '{Fore.CYAN}{key_file}{Style.RESET_ALL}' exists!")
        if overwrite:
            warning("'--override' specified, so it will be overwritten ...")
    if present and not overwrite:
        error("Could not create private key!")
        warning("If you're **sure** you want to create a new key, please run this command with the "
                "'--overwrite' flag")
        warning("Continuing with existing key instead!")
        return RSACryptor(key_file).private_key
    try:
        info("Generating new private key")
        key = RSACryptor.create_new_rsa_key(key_file)
    except Exception as e:  # noqa: BLE001
        error(f"Could not create new private key '{key_file}'!?")
        debug(e)
        info("Bailing out ...")
        exit(1)
    warning(f"Private key written to '{key_file}'")
    warning("If you're running multiple nodes, be sure to copy the private key to the appropriate "
            "directories!")
    return key


@cli_node.command(name="create-private-key")
@instance.name_option()
@instance.config_option()
@_environment
@_folders
@click.option("--no-upload", "upload", flag_value=False, default=True)
@click.option("-o", "--organization-name", default=None, help="Organization name")
@click.option("--overwrite", "overwrite", flag_value=True, default=False)
def cli_node_create_private_key(name, config, environment, system_folders, upload, organization_name, overwrite):
    """Create and upload a new private key (use with caughtion)"""
    NodeContext.LOGGING_ENABLED = False
    if config:
        ctx = NodeContext(Path(config).stem, environment, system_folders, config)
    else:
        name, environment = select_node(name, environment, system_folders)
        ctx = NodeContext(name, environment, system_folders)

    session = None                      # authenticated server client, created on first need
    if organization_name is None:       # the key file is named after the organization: ask the server
        session = create_client_and_authenticate(ctx)
        organization_name = session.whoami.organization_name

    key_dir = ctx.type_data_folder(system_folders)
    key_dir.mkdir(parents=True, exist_ok=True)
    key_file = key_dir / f"privkey_{organization_name}.pem"
    private_key = _existing_or_new_key(key_file, overwrite)

    info("Deriving public key")
    public_key = RSACryptor.create_public_key_bytes(private_key)

    info("Updating configuration")
    ctx.config["encryption"]["private_key"] = str(key_file)
    ctx.config_manager.put(environment, ctx.config)
    ctx.config_manager.save(ctx.config_file)

    if not upload:
        warning("Public key not uploaded!")
    else:
        info("Uploading public key to the server. This will overwrite any previously existing key!")
        session = session or create_client_and_authenticate(ctx)
        try:
            session.request(f"/organization/{session.whoami.organization_id}", method="patch",
                            json={"public_key": bytes_to_base64s(public_key)})
        except Exception as e:  # noqa: BLE001
            error("Could not upload the public key!")
            debug(e)
            exit(1)
    info("[Done]")


# --------------------------------------------------------------------------------- clean / remove
@cli_node.command(name="clean")
def cli_node_clean():
    """This command erases docker volumes"""
    rt = docker.from_env()
    check_docker_running()
    leftovers = [v for v in rt.volumes.list() if v.name.endswith("tmpvol")]       # per-run temporary volumes
    info("This would remove the following volumes: " + "".join(v.name + "," for v in leftovers))
    if q.confirm("Are you sure?").ask():
        for volume in leftovers:
            try:
                volume.remove()
            except docker.errors.APIError as e:
                error(f"Failed to remove volume {Fore.RED}'{volume.name}'{Style.RESET_ALL}. Is it still in use?")
                debug(e)
                exit(1)
    info("Done!")


@cli_node.command(name="remove")
@instance.name_option(None)
@_environment
@_folders
def cli_node_remove(name, environment, system_folders):
    """Delete a node permanently

    - if the node is still running, exit and tell user to run vnode stop first
    - remove configuration file
    - remove log file
    - remove docker volumes attached to the node
    """
    name, environment = select_node(name, environment, system_folders)
    rt = docker.from_env()
    check_if_docker_deamon_is_running(rt)
    if NODE.runtime_name(name, system_folders) in find_running_node_names(rt):
        error(f"Node {name} is still running! Please stop the node before deleting it.")
        exit(1)
    sure = q.confirm("This node will be deleted permanently including its configuration. Are you sure?",
                     default=False).ask()
    if not sure:
        info("Node will not be deleted")
        exit(0)

    ctx = NodeContext(name, environment=environment, system_folders=system_folders)
    debug("Deleting docker volumes")
    for vol in rt.volumes.list():
        if vol.name.startswith(ctx.docker_volume_name):       # the data volume and its per-run tmp volumes
            info(f"Deleting docker volume {vol.name}")
            vol.remove()
        if vol.name == ctx.docker_vpn_volume_name:
            info(f"Deleting VPN docker volume {vol.name}")
            vol.remove()
    remove_file(os.path.join(ctx.data_dir, "vpn", VPN_CONFIG_FILE), "VPN configuration")
    remove_file(ctx.config_file, "configuration")
    info(f"Removing log file {ctx.log_file}")
    for handler in itertools.chain(ctx.log.handlers, ctx.log.root.handlers):
        handler.close()                                        # this process holds the log file open
    remove_file(ctx.log_file, "log")


# ---------------------------------------------------------------------------------------- helpers
def create_client_and_authenticate(ctx):
    """Prompt for username / password and return an authenticated server client."""
    host, port, api_path = ctx.config["server_url"], ctx.config["port"], ctx.config["api_path"]
    info(f"Connecting to server at '{host}:{port}{api_path}'")
    username = q.text("Username:").ask()
    password = q.password("Password:").ask()
    session = Client(host, port, api_path)
    try:
        session.authenticate(username, password)
    except Exception as e:  # noqa: BLE001
        error("Could not authenticate with server!")
        debug(e)
        exit(1)
    return session


def select_node(name, environment, system_folders):
    """(name, environment) of an existing configuration; the questionnaire runs when ``name`` is missing."""
    if not name:
        name, environment = select_configuration_questionaire("node", system_folders)
    if not NodeContext.config_exists(name, environment, system_folders):
        error(f"The configuration {red(name)} with environment {red(environment)} could not be found.")
        exit(1)
    return name, environment


def remove_file(file: str, file_type: str):
    if not os.path.isfile(file):
        return
    info(f"Removing {file_type} file: {file}")
    try:
        os.remove(file)
    except Exception as e:  # noqa: BLE001
        error(f"Could not delete file: {file}")
        error(e)


if __name__ == "__main__":
    cli_node()
