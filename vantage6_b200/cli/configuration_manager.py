"""Configuration schemas of server and node instances
(reference vantage6/cli/configuration_manager.py:9-92; Appendix B of SURVEY.md).

Optional keys tolerated by the schema: ``image``, ``vpn_subnet`` / ``vpn_server``,
``jwt_secret_key``, ``rabbitmq_uri`` plus the B200 extensions ``gpu`` (node: device index) and
``algorithms`` (node: image -> python module map)."""
from ..common.configuration_manager import Configuration, ConfigurationManager
from ..common.schema import And, Optional, Or, Use

_LOGGING = {
    "level": And(Use(str), lambda lvl: lvl in ("DEBUG", "INFO", "WARNING", "ERROR", "CRITICAL", "NOTSET")),
    "file": Use(str),
    "use_console": Use(bool),
    "backup_count": And(Use(int), lambda n: n > 0),
    "max_size": And(Use(int), lambda b: b > 16),
    "format": Use(str),
    "datefmt": Use(str),
}


class ServerConfiguration(Configuration):
    VALIDATORS = {
        "description": Use(str),
        "ip": Use(str),
        "port": Use(int),
        "api_path": Use(str),
        "uri": Use(str),
        "allow_drop_all": Use(bool),
        "logging": dict(_LOGGING),
    }


class NodeConfiguration(Configuration):
    VALIDATORS = {
        "api_key": And(Use(str), len),
        "server_url": Use(str),
        "port": Or(Use(int), None),
        "task_dir": Use(str),
        "databases": {Use(str): Use(str)},
        "api_path": Use(str),
        "logging": dict(_LOGGING),
        "encryption": {"enabled": bool, Optional("private_key"): Use(str)},
    }


class TestConfiguration(Configuration):
    VALIDATORS = {}


class NodeConfigurationManager(ConfigurationManager):
    def __init__(self, name, *args, **kwargs):
        super().__init__(conf_class=NodeConfiguration, name=name)

    @classmethod
    def from_file(cls, path):
        return super().from_file(path, conf_class=NodeConfiguration)


class ServerConfigurationManager(ConfigurationManager):
    def __init__(self, name, *args, **kwargs):
        super().__init__(conf_class=ServerConfiguration, name=name)

    @classmethod
    def from_file(cls, path):
        return super().from_file(path, conf_class=ServerConfiguration)
