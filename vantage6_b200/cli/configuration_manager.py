"""What a valid server / node configuration looks like, and the managers that load and save them.

The validators are assembled from small field helpers (``_text``, ``_integer``, ...) instead of being spelled out
per key; the resulting schemas accept exactly what the reference accepts
(vantage6/cli/configuration_manager.py:9-92).  Optional keys tolerated on top:
``image``, ``vpn_subnet`` / ``vpn_server``, ``jwt_secret_key``, ``rabbitmq_uri`` and the B200 extensions ``gpu``
(node: device index) and ``algorithms`` (node: image name -> python module).
"""
from ..common.configuration_manager import Configuration, ConfigurationManager
from ..common.schema import And, Optional, Or, Use

_LOG_LEVELS = ("DEBUG", "INFO", "WARNING", "ERROR", "CRITICAL", "NOTSET")


def _text(non_empty: bool = False):
    return And(Use(str), len) if non_empty else Use(str)


def _integer(above=None):
    return Use(int) if above is None else And(Use(int), lambda value: value > above)


def _logging_block() -> dict:
    return {"level": And(Use(str), lambda level: level in _LOG_LEVELS), "file": _text(), "use_console": Use(bool),
            "backup_count": _integer(above=0), "max_size": _integer(above=16), "format": _text(), "datefmt": _text()}


def _schema(**fields) -> dict:
    fields["logging"] = _logging_block()
    return fields


class ServerConfiguration(Configuration):
    VALIDATORS = _schema(description=_text(), ip=_text(), port=_integer(), api_path=_text(), uri=_text(),
                         allow_drop_all=Use(bool))


class NodeConfiguration(Configuration):
    VALIDATORS = _schema(api_key=_text(non_empty=True), server_url=_text(), port=Or(Use(int), None), task_dir=_text(),
                         api_path=_text(), databases={Use(str): Use(str)},
                         encryption={"enabled": bool, Optional("private_key"): Use(str)})


class TestConfiguration(Configuration):
    VALIDATORS = {}


def _manager_for(conf_class, doc: str):
    """A ``ConfigurationManager`` subclass bound to one configuration class."""

    class _Manager(ConfigurationManager):
        __doc__ = doc

        def __init__(self, name, *args, **kwargs):
            super().__init__(conf_class=conf_class, name=name)

        @classmethod
        def from_file(cls, path):
            return super().from_file(path, conf_class=conf_class)

    return _Manager


NodeConfigurationManager = _manager_for(NodeConfiguration, "Loads / saves multi-environment node configuration files.")
NodeConfigurationManager.__name__ = NodeConfigurationManager.__qualname__ = "NodeConfigurationManager"
ServerConfigurationManager = _manager_for(ServerConfiguration, "Loads / saves multi-environment server configuration files.")
ServerConfigurationManager.__name__ = ServerConfigurationManager.__qualname__ = "ServerConfigurationManager"
