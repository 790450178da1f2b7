"""``vnode`` -- node manager command line interface.

Behavioural spec: reference vantage6/cli/node.py (commands ``list new files start stop attach
create-private-key clean remove version`` with the same options, defaults, console messages and
exit codes).  What differs is *what gets launched*: a federated node is
one process pinned to one B200 (``--gpu K`` -> ``V6_GPU=K``) managed by the
process runtime (vantage6_b200/runtime) instead of a Docker container.
"""
from __future__ import annotations

import itertools
import os.path
import sys
import time
from pathlib import Path
from threading import Thread

import click

from .. import runtime as docker            # docker-SDK-shaped process runtime
from ..client import Client
from ..client.encryption import RSACryptor
from ..common import bytes_to_base64s, check_config_write_permissions, debug, error, info, warning
from ..common import prompts as q
from ..common.colors import Fore, Style
from ..common.globals import APPNAME, DEFAULT_DOCKER_REGISTRY, DEFAULT_NODE_IMAGE, STRING_ENCODING, VPN_CONFIG_FILE
from ..runtime.addons import check_docker_running, pull_if_newer, remove_container_if_exists
from . import __version__
from .configuration_wizard import configuration_wizard, select_configuration_questionaire
from .context import NodeContext
from .globals import DEFAULT_NODE_ENVIRONMENT as N_ENV
from .globals import DEFAULT_NODE_SYSTEM_FOLDERS as N_FOL
from .utils import check_config_name_allowed, check_if_docker_deamon_is_running

LIST_HEADER = "\nName" + (21 * " ") + "Environments" + (20 * " ") + "Status" + (10 * " ") + "System/User"


@click.group(name="node")
def cli_node():
    """Subcommand `vnode`."""


def _scope(system_folders: bool) -> str:
    return "system" if system_folders else "user"


def _system_user_options(fn):
    fn = click.option("--user", "system_folders", flag_value=False, default=N_FOL)(fn)
    return click.option("--system", "system_folders", flag_value=True)(fn)


def find_running_node_names(client):
    running_nodes = client.containers.list(filters={"label": f"{APPNAME}-type=node"})
    return [node.name for node in running_nodes]


def print_log_worker(logs_stream):
    for log in logs_stream:
        print(log.decode(STRING_ENCODING), end="")


def _follow_logs(container):
    """Stream a running node's log until Ctrl-C (reference node.py:414-422, 491-500)."""
    logs = container.attach(stream=True, logs=True)
    Thread(target=print_log_worker, args=(logs,), daemon=True).start()
    while True:
        try:
            time.sleep(1)
        except KeyboardInterrupt:
            info("Closing log file. Keyboard Interrupt.")
            exit(0)


# --------------------------------------------------------------------------------------- list
@cli_node.command(name="list")
def cli_node_list():
    """Lists all nodes in the default configuration directory."""
    client = docker.from_env()
    check_docker_running()
    running_node_names = find_running_node_names(client)

    click.echo(LIST_HEADER)
    click.echo("-" * len(LIST_HEADER))
    running = Fore.GREEN + "Online" + Style.RESET_ALL
    stopped = Fore.RED + "Offline" + Style.RESET_ALL
    failed = 0
    for system_folders, tag in ((True, " System "), (False, " User   ")):
        configs, f = NodeContext.available_configurations(system_folders=system_folders)
        failed += len(f)
        for config in configs:
            online = f"{APPNAME}-{config.name}-{_scope(system_folders)}" in running_node_names
            status = running if online else stopped
            click.echo(f"{config.name:25}{str(config.available_environments):32}{status:25}{tag}")
    click.echo("-" * 85)
    if failed:
        warning(f"{Fore.RED}Failed imports: {failed}{Style.RESET_ALL}")


# ---------------------------------------------------------------------------------------- new
@cli_node.command(name="new")
@click.option("-n", "--name", default=None)
@click.option("-e", "--environment", default="", help="configuration environment to use")
@_system_user_options
def cli_node_new_configuration(name, environment, system_folders):
    """Create a new configuration file.

    Checks if the configuration already exists. If this is not the case a questionaire is
    invoked to create a new configuration file.
    """
    if not name:
        name = q.text("Please enter a configuration-name:").ask()
    name_new = name.replace(" ", "-")
    if name != name_new:
        info(f"Replaced spaces from configuration name: {name_new}")
        name = name_new
    check_config_name_allowed(name)

    if not environment:
        environment = q.select("Please select the environment you want to configure:",
                               ["application", "prod", "acc", "test", "dev"]).ask()

    if NodeContext.config_exists(name, environment, system_folders):
        error(f"Configuration {name} and environment{environment} already exists!")
        exit(1)
    if not check_config_write_permissions(system_folders):
        error("Your user does not have write access to all folders. Exiting")
        exit(1)

    flag = "--system" if system_folders else ""
    cfg_file = configuration_wizard("node", name, environment, system_folders)
    info(f"New configuration created: {Fore.GREEN}{cfg_file}{Style.RESET_ALL}")
    info(f"You can start the node by running {Fore.GREEN}vnode start {flag}{Style.RESET_ALL}")


# -------------------------------------------------------------------------------------- files
@cli_node.command(name="files")
@click.option("-n", "--name", default=None, help="configuration name")
@click.option("-e", "--environment", default=N_ENV, help="configuration environment to use")
@_system_user_options
def cli_node_files(name, environment, system_folders):
    """Prints location important files.

    If the specified configuration cannot be found, it exits. Otherwise it returns the
    absolute path to the output.
    """
    name, environment = select_node(name, environment, system_folders)
    ctx = NodeContext(name, environment=environment, system_folders=system_folders)
    info(f"Configuration file = {ctx.config_file}")
    info(f"Log file           = {ctx.log_file}")
    info(f"data folders       = {ctx.data_dir}")
    info("Database labels and files")
    for label, path in ctx.databases.items():
        info(f" - {label:15} = {path}")


# -------------------------------------------------------------------------------------- start
@cli_node.command(name="start")
@click.option("-n", "--name", default=None, help="configuration name")
@click.option("-c", "--config", default=None, help="absolute path to configuration-file; overrides NAME")
@click.option("-e", "--environment", default=N_ENV, help="configuration environment to use")
@_system_user_options
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
    docker_client = docker.from_env()
    check_docker_running()

    NodeContext.LOGGING_ENABLED = False
    if config:
        name = Path(config).stem
        ctx = NodeContext(name, environment, system_folders, config)
    else:
        if not name:
            name, environment = select_configuration_questionaire("node", system_folders)
        if not NodeContext.config_exists(name, environment, system_folders):
            question = (f"Configuration '{name}' using environment '{environment}' does not exist.\n"
                        "  Do you want to create this config now?")
            if q.confirm(question).ask():
                configuration_wizard("node", name, environment, system_folders)
            else:
                error("Config file couldn't be loaded")
                sys.exit(0)
        ctx = NodeContext(name, environment, system_folders)

    check_config_name_allowed(ctx.name)

    # refuse to start the same node twice
    suffix = _scope(system_folders)
    for node in docker_client.containers.list(filters={"label": f"{APPNAME}-type=node"}):
        if node.name == f"{APPNAME}-{name}-{suffix}":
            error(f"Node {Fore.RED}{name}{Style.RESET_ALL} is already running")
            exit(1)

    info("Checking that data and log dirs exist")
    ctx.data_dir.mkdir(parents=True, exist_ok=True)
    ctx.log_dir.mkdir(parents=True, exist_ok=True)

    # image precedence: CLI option > configuration file > package default
    if not image:
        image = ctx.config.get("image", f"{DEFAULT_DOCKER_REGISTRY}/{DEFAULT_NODE_IMAGE}")
    info(f"Pulling latest node image '{image}'")
    try:
        pull_if_newer(docker.from_env(), image)
    except Exception:  # noqa: BLE001
        warning(" ... alas, no dice!")
    else:
        info(" ... success!")

    info("Creating Docker data volume")
    data_volume = docker_client.volumes.create(ctx.docker_volume_name)
    vpn_volume = docker_client.volumes.create(ctx.docker_vpn_volume_name)

    info("Creating file & folder mounts")
    mounts = [
        # (target, source)
        ("/mnt/log", str(ctx.log_dir)),
        ("/mnt/data", data_volume.name),
        ("/mnt/vpn", vpn_volume.name),
        ("/mnt/config", str(ctx.config_dir)),
    ]
    if mount_src:
        mounts.append(("/vantage6", os.path.abspath(mount_src)))

    # private key: config -> default name -> PRIVATE_KEY environment override
    filename = ctx.config.get("encryption", {}).get("private_key")
    if not filename:
        filename = "private_key.pem"
    filename = os.environ.get("PRIVATE_KEY", filename)
    fullpath = Path(ctx.get_data_file(filename))
    env = {"DATA_VOLUME_NAME": data_volume.name, "VPN_VOLUME_NAME": vpn_volume.name}
    if fullpath:
        if Path(fullpath).exists():
            mounts.append(("/mnt/private_key.pem", str(fullpath)))
            env["PRIVATE_KEY"] = str(fullpath)
        else:
            warning(f"private key file provided {fullpath}, but does not exists")

    info("Setting up databases")
    for label in ctx.databases.keys():
        uri = ctx.databases[label]
        info(f"  Processing database '{label}:{uri}'")
        LABEL = label.upper()
        file_based = Path(uri).exists()
        if not file_based and not force_db_mount:
            debug("  - non file-based database added")
            env[f"{LABEL}_DATABASE_URI"] = uri
        else:
            debug("  - file-based database added")
            env[f"{LABEL}_DATABASE_URI"] = str(Path(uri).resolve())
            mounts.append((f"/mnt/{label}.csv", str(uri)))
        if label == "default":
            env["DATABASE_URI"] = env[f"{LABEL}_DATABASE_URI"]

    if gpu is None and ctx.config.get("gpu") is not None:
        gpu = int(ctx.config.get("gpu"))
    if gpu is not None:
        env["V6_GPU"] = str(gpu)              # all GPUs stay visible: peers are mapped over NVLink (symmetric heap)
        info(f"Pinning node to GPU {gpu}")

    system_folders_option = "--system" if system_folders else "--user"
    cmd = (f"vnode-local start -c /mnt/config/{name}.yaml -n {name} -e {environment} "
           f"--dockerized {system_folders_option}")

    info("Running Docker container")
    volumes = [f"{source}:{target}" for (target, source) in mounts]
    remove_container_if_exists(docker_client=docker_client, name=ctx.docker_container_name)
    container = docker_client.containers.run(
        image, command=cmd, volumes=volumes, detach=True,
        labels={f"{APPNAME}-type": "node", "system": str(system_folders), "name": ctx.config_file_name},
        environment=env, name=ctx.docker_container_name, auto_remove=not keep, tty=True)
    info(f"Success! container id = {container}")

    if attach:
        _follow_logs(container)


# --------------------------------------------------------------------------------------- stop
@cli_node.command(name="stop")
@click.option("-n", "--name", default=None, help="configuration name")
@_system_user_options
@click.option("--all", "all_nodes", flag_value=True)
def cli_node_stop(name, system_folders, all_nodes):
    """Stop a running container."""
    client = docker.from_env()
    check_docker_running()
    running_node_names = find_running_node_names(client)
    if not running_node_names:
        warning("No nodes are currently running.")
        return

    if all_nodes:
        for name in running_node_names:
            client.containers.get(name).kill()
            info(f"Stopped the {Fore.GREEN}{name}{Style.RESET_ALL} Node.")
        return

    if not name:
        name = q.select("Select the node you wish to stop:", choices=running_node_names).ask()
    else:
        name = f"{APPNAME}-{name}-{_scope(system_folders)}"
    if name in running_node_names:
        # stop() gives the node 10 s to exit by itself, then it is killed
        client.containers.get(name).stop()
        info(f"Stopped the {Fore.GREEN}{name}{Style.RESET_ALL} Node.")
    else:
        error(f"{Fore.RED}{name}{Style.RESET_ALL} is not running?")


# ------------------------------------------------------------------------------------- attach
@cli_node.command(name="attach")
@click.option("-n", "--name", default=None, help="configuration name")
@_system_user_options
def cli_node_attach(name, system_folders):
    """Attach the logs from the docker container to the terminal."""
    client = docker.from_env()
    check_docker_running()
    running_node_names = find_running_node_names(client)
    if not name:
        name = q.select("Select the node you wish to inspect:", choices=running_node_names).ask()
    else:
        name = f"{APPNAME}-{name}-{_scope(system_folders)}"
    if name in running_node_names:
        _follow_logs(client.containers.get(name))
    else:
        error(f"{Fore.RED}{name}{Style.RESET_ALL} was not running!?")


# ------------------------------------------------------------------------- create-private-key
@cli_node.command(name="create-private-key")
@click.option("-n", "--name", default=None, help="configuration name")
@click.option("-c", "--config", default=None, help="absolute path to configuration-file; overrides NAME")
@click.option("-e", "--environment", default=N_ENV, help="configuration environment to use")
@_system_user_options
@click.option("--no-upload", "upload", flag_value=False, default=True)
@click.option("-o", "--organization-name", default=None, help="Organization name")
@click.option("--overwrite", "overwrite", flag_value=True, default=False)
def cli_node_create_private_key(name, config, environment, system_folders, upload, organization_name, overwrite):
    """Create and upload a new private key (use with caughtion)"""
    NodeContext.LOGGING_ENABLED = False
    if config:
        name = Path(config).stem
        ctx = NodeContext(name, environment, system_folders, config)
    else:
        name, environment = select_node(name, environment, system_folders)
        if not NodeContext.config_exists(name, environment, system_folders):
            error(f"The configuration {Fore.RED}{name}{Style.RESET_ALL} with environment "
                  f"{Fore.RED}{environment}{Style.RESET_ALL} could not be found.")
            exit(1)
        ctx = NodeContext(name, environment, system_folders)

    # the organization name (needed for the key's file name) comes from the server if not given
    client = None
    if organization_name is None:
        client = create_client_and_authenticate(ctx)
        organization_name = client.whoami.organization_name

    ctx.type_data_folder(system_folders).mkdir(parents=True, exist_ok=True)
    filename = f"privkey_{organization_name}.pem"
    file_ = ctx.type_data_folder(system_folders) / filename

    if file_.exists():
        warning(f"File '{Fore.CYAN}{file_}{Style.RESET_ALL}' exists!")
        if overwrite:
            warning("'--override' specified, so it will be overwritten ...")

    if file_.exists() and not overwrite:
        error("Could not create private key!")
        warning("If you're **sure** you want to create a new key, please run this command with the "
                "'--overwrite' flag")
        warning("Continuing with existing key instead!")
        private_key = RSACryptor(file_).private_key
    else:
        try:
            info("Generating new private key")
            private_key = RSACryptor.create_new_rsa_key(file_)
        except Exception as e:  # noqa: BLE001
            error(f"Could not create new private key '{file_}'!?")
            debug(e)
            info("Bailing out ...")
            exit(1)
        warning(f"Private key written to '{file_}'")
        warning("If you're running multiple nodes, be sure to copy the private key to the appropriate "
                "directories!")

    info("Deriving public key")
    public_key = RSACryptor.create_public_key_bytes(private_key)

    info("Updating configuration")
    ctx.config["encryption"]["private_key"] = str(file_)
    ctx.config_manager.put(environment, ctx.config)
    ctx.config_manager.save(ctx.config_file)

    if upload:
        info("Uploading public key to the server. This will overwrite any previously existing key!")
        if client is None:
            client = create_client_and_authenticate(ctx)
        try:
            client.request(f"/organization/{client.whoami.organization_id}", method="patch",
                           json={"public_key": bytes_to_base64s(public_key)})
        except Exception as e:  # noqa: BLE001
            error("Could not upload the public key!")
            debug(e)
            exit(1)
    else:
        warning("Public key not uploaded!")
    info("[Done]")


# -------------------------------------------------------------------------------------- clean
@cli_node.command(name="clean")
def cli_node_clean():
    """This command erases docker volumes"""
    client = docker.from_env()
    check_docker_running()

    canditates = []
    msg = "This would remove the following volumes: "
    for volume in client.volumes.list():
        if volume.name[-6:] == "tmpvol":
            canditates.append(volume)
            msg += volume.name + ","
    info(msg)

    if q.confirm("Are you sure?").ask():
        for volume in canditates:
            try:
                volume.remove()
            except docker.errors.APIError as e:
                error(f"Failed to remove volume {Fore.RED}'{volume.name}'{Style.RESET_ALL}. Is it still in use?")
                debug(e)
                exit(1)
    info("Done!")


# ------------------------------------------------------------------------------------- remove
@cli_node.command(name="remove")
@click.option("-n", "--name", default=None)
@click.option("-e", "--environment", default=N_ENV, help="configuration environment to use")
@_system_user_options
def cli_node_remove(name, environment, system_folders):
    """Delete a node permanently

    - if the node is still running, exit and tell user to run vnode stop first
    - remove configuration file
    - remove log file
    - remove docker volumes attached to the node
    """
    name, environment = select_node(name, environment, system_folders)
    client = docker.from_env()
    check_if_docker_deamon_is_running(client)

    node_container_name = f"{APPNAME}-{name}-{_scope(system_folders)}"
    if node_container_name in find_running_node_names(client):
        error(f"Node {name} is still running! Please stop the node before deleting it.")
        exit(1)

    if not q.confirm("This node will be deleted permanently including its configuration. Are you sure?",
                     default=False).ask():
        info("Node will not be deleted")
        exit(0)

    ctx = NodeContext(name, environment=environment, system_folders=system_folders)

    debug("Deleting docker volumes")
    for vol in client.volumes.list():
        if vol.name.startswith(ctx.docker_volume_name):      # includes the per-run tmp volumes
            info(f"Deleting docker volume {vol.name}")
            vol.remove()
        if vol.name == ctx.docker_vpn_volume_name:
            info(f"Deleting VPN docker volume {vol.name}")
            vol.remove()

    remove_file(os.path.join(ctx.data_dir, "vpn", VPN_CONFIG_FILE), "VPN configuration")
    remove_file(ctx.config_file, "configuration")

    # this process opened the log file above: close the handlers before deleting it
    info(f"Removing log file {ctx.log_file}")
    for handler in itertools.chain(ctx.log.handlers, ctx.log.root.handlers):
        handler.close()
    remove_file(ctx.log_file, "log")


# ------------------------------------------------------------------------------------ version
@cli_node.command(name="version")
@click.option("-n", "--name", default=None, help="configuration name")
@_system_user_options
def cli_node_version(name, system_folders):
    """Returns current version of vantage6 services installed."""
    client = docker.from_env()
    check_docker_running()
    running_node_names = find_running_node_names(client)

    if not name:
        if not running_node_names:
            error("No nodes are running! You can only check the version for nodes that are running")
            exit(1)
        name = q.select("Select the node you wish to inspect:", choices=running_node_names).ask()
    else:
        name = f"{APPNAME}-{name}-{_scope(system_folders)}"

    if name in running_node_names:
        container = client.containers.get(name)
        version = container.exec_run(cmd="vnode-local version", stdout=True)
        click.echo({"node": version.output.decode("utf-8"), "cli": __version__})
    else:
        error(f"Node {name} is not running! Cannot provide version...")


# ------------------------------------------------------------------------------------ helpers
def create_client_and_authenticate(ctx):
    """Create a client and authenticate (username/password prompt)."""
    host = ctx.config["server_url"]
    port = ctx.config["port"]
    api_path = ctx.config["api_path"]

    info(f"Connecting to server at '{host}:{port}{api_path}'")
    username = q.text("Username:").ask()
    password = q.password("Password:").ask()

    client = Client(host, port, api_path)
    try:
        client.authenticate(username, password)
    except Exception as e:  # noqa: BLE001
        error("Could not authenticate with server!")
        debug(e)
        exit(1)
    return client


def select_node(name, environment, system_folders):
    """Let the user pick a configuration through the questionnaire if ``name`` is unknown."""
    name, environment = (name, environment) if name else select_configuration_questionaire("node", system_folders)
    if not NodeContext.config_exists(name, environment, system_folders):
        error(f"The configuration {Fore.RED}{name}{Style.RESET_ALL} with environment "
              f"{Fore.RED}{environment}{Style.RESET_ALL} could not be found.")
        exit(1)
    return name, environment


def remove_file(file: str, file_type: str):
    if os.path.isfile(file):
        info(f"Removing {file_type} file: {file}")
        try:
            os.remove(file)
        except Exception as e:  # noqa: BLE001
            error(f"Could not delete file: {file}")
            error(e)
    else:
        warning(f"Could not remove {file_type} file: {file} does not exist")


if __name__ == "__main__":
    cli_node()
