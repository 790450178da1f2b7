"""Interactive configuration wizard.

The questions are data: each questionnaire is a table of ``(key, prompt, default)`` rows (plus a few
follow-up blocks guarded by a yes/no question), turned into prompt specifications by :func:`_ask`.  Questions,
defaults, order and resulting keys are the reference's (vantage6/cli/configuration_wizard.py:13-273), so answer
scripts written for it keep working; port / ``allow_drop_all`` answers are coerced to the schema's types before
the file is written.

Entry points: :func:`configuration_wizard` (run a questionnaire, write or extend ``<config dir>/<name>.yaml``),
:func:`select_configuration_questionaire` (pick one of the existing configurations),
:func:`node_configuration_questionaire`, :func:`server_configuration_questionaire`.
"""
from __future__ import annotations

import uuid
from pathlib import Path

from ..common import prompts as q
from .configuration_manager import NodeConfigurationManager, ServerConfigurationManager
from .context import NodeContext, ServerContext

LOG_LEVELS = ["DEBUG", "INFO", "WARNING", "ERROR", "CRITICAL", "NOTSET"]
_LOG_FORMAT = {"format": "%(asctime)s - %(name)-14s - %(levelname)-8s - %(message)s", "datefmt": "%Y-%m-%d %H:%M:%S"}

# (key, prompt, default or None[, prompt type])
_NODE_BASICS = (
    ("api_key", "Enter given api-key:", None),
    ("server_url", "The base-URL of the server:", "http://localhost"),
    ("port", "Enter port to which the server listens:", "5000"),
    ("api_path", "Path of the api:", "/api"),
)
_SERVER_BASICS = (
    ("description", "Enter a human-readable description:", None),
    ("ip", "ip:", "0.0.0.0"),
    ("port", "Enter port to which the server listens:", "5000"),
    ("api_path", "Path of the api:", "/api"),
    ("uri", "Database URI:", "sqlite:///default.sqlite"),
)
_VPN_SERVER = (
    ("url", "VPN server URL:", None),
    ("portal_username", "VPN portal username:", None),
    ("portal_userpass", "VPN portal password:", None, "password"),
    ("client_id", "VPN client username:", None),
    ("client_secret", "VPN client password:", None, "password"),
    ("redirect_url", "Redirect url (should be local address of server)", "http://localhost"),
)


def _spec(key, message, default=None, kind="text") -> dict:
    spec = {"type": kind, "name": key, "message": message}
    if default is not None:
        spec["default"] = default
    return spec


def _ask(rows, *extra) -> dict:
    """Prompt for a table of rows (and any ready-made specifications in ``extra``) in one go."""
    return q.prompt([_spec(*row) for row in rows] + list(extra))


def _yes(question: str, **kw) -> bool:
    return bool(q.confirm(question, **kw).ask())


def _logging_section(instance_name: str, level: str) -> dict:
    """Not asked, fixed: rotating file ``<instance>.log`` (5 backups of 1024 kB) plus the console."""
    return {"level": level, "file": f"{instance_name}.log", "use_console": True, "backup_count": 5, "max_size": 1024,
            **_LOG_FORMAT}


def _ask_log_level() -> str:
    return q.select("Which level of logging would you like?", choices=LOG_LEVELS).ask()


def node_configuration_questionaire(dirs, instance_name):
    """Everything a node needs to know: its server, its task directory, its labelled databases, logging,
    optionally a VPN subnet, and whether task payloads are encrypted."""
    config = _ask(_NODE_BASICS + (("task_dir", "Task directory path:", str(dirs["data"])),))

    databases = _ask((("default", "Default database path:", None),))
    config["databases"] = databases
    extra_index = 1
    while _yes("Do you want to add another database?"):
        near_default = str(Path(databases.get("default")).parent)
        added = _ask((("label", "Enter the label for the database:", f"database_{extra_index}"),
                      ("path", "The path of the database file:", near_default)))
        databases[added.get("label")] = added.get("path")
        extra_index += 1

    level = _ask_log_level()
    if _yes("Do you want to connect to a VPN server?", default=False):
        config["vpn_subnet"] = q.text(message="Subnet of the VPN server you want to connect to:",
                                      default="10.76.0.0/16").ask()
    config["logging"] = _logging_section(instance_name, level)

    encrypted = q.select("Enable encryption?", choices=["true", "false"]).ask() == "true"
    config["encryption"] = {"enabled": encrypted,
                            "private_key": q.text("Path to private key file:").ask() if encrypted else ""}
    return config


def server_configuration_questionaire(dirs, instance_name):
    """Everything a server needs to know: where it listens, its database, whether ``drop all`` is allowed, and the
    optional extras (constant JWT secret, VPN server, message queue)."""
    drop_all = {"type": "select", "name": "allow_drop_all", "message": "Allowed to drop all tables: ",
                "choices": ["True", "False"]}
    config = _ask(_SERVER_BASICS, drop_all)

    if _yes("Do you want a constant JWT secret?"):
        config["jwt_secret_key"] = str(uuid.uuid1())
    level = _ask_log_level()
    if _yes("Do you want to add a VPN server?", default=False):
        config["vpn_server"] = _ask(_VPN_SERVER)
    if _yes("Do you want to add a RabbitMQ message queue?"):
        config["rabbitmq_uri"] = q.text(message="Enter the URI for your RabbitMQ:").ask()
    config["logging"] = _logging_section(instance_name, level)
    return config


def _coerce_answers(type_: str, config: dict) -> dict:
    """Prompt answers are strings; the schema wants ``port`` as int and ``allow_drop_all`` as bool."""
    port = config.get("port")
    if port not in (None, ""):
        try:
            config["port"] = int(port)
        except (TypeError, ValueError):
            pass                                     # left for the schema to complain about
    flag = config.get("allow_drop_all")
    if type_ == "server" and isinstance(flag, str):
        config["allow_drop_all"] = flag.strip().lower() == "true"
    return config


def configuration_wizard(type_, instance_name, environment, system_folders):
    """Run the questionnaire for ``type_`` and store the answers as ``environment`` of
    ``<config dir>/<instance_name>.yaml``.  A file that already exists is extended (or that environment replaced),
    never overwritten as a whole.  Returns the path."""
    folders = NodeContext.instance_folders(type_, instance_name, system_folders)
    questionnaire, manager_class = ((node_configuration_questionaire, NodeConfigurationManager) if type_ == "node"
                                    else (server_configuration_questionaire, ServerConfigurationManager))
    answers = questionnaire(folders, instance_name)
    if isinstance(answers, dict):
        answers = _coerce_answers(type_, answers)

    target = Path(folders.get("config")) / f"{instance_name}.yaml"
    manager = manager_class.from_file(target) if target.exists() else manager_class(instance_name)
    manager.put(environment, answers)
    manager.save(target)
    return target


def select_configuration_questionaire(type_, system_folders):
    """Let the user pick one (configuration, environment) pair from the default folder of ``type_``."""
    context = NodeContext if type_ == "node" else ServerContext
    configs, _broken = context.available_configurations(system_folders)
    menu = [q.Choice(title=f"{cfg.name:25} {env}", value=(cfg.name, env))
            for cfg in configs for env in cfg.available_environments]
    if not menu:
        raise Exception("No configurations could be found!")
    name, env = q.select("Select the configuration you want to use:", choices=menu).ask()
    return name, env
