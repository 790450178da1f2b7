"""Server / node application contexts.

Both are an :class:`~vantage6_b200.common.context.AppContext` bound to an instance type; the binding
(type name, default environment, default folder scope) is declared once per class and the
class-level queries (``config_exists``, ``available_configurations``, ``from_external_config_file``)
are implemented once in :class:`_BoundContext`.  On top of that the contexts know the names the
runtime uses for an instance -- process ("container"), network, data / VPN / per-run temporary
volumes -- and where its database lives, including the ``VANTAGE6_DB_URI`` / ``VANTAGE6_CONFIG_NAME``
/ ``DATA_VOLUME_NAME`` / ``VPN_VOLUME_NAME`` environment overrides.

Naming scheme and override semantics follow the reference (vantage6/cli/context.py:16-144).
"""
import os
from urllib.parse import urlsplit

from .._version import __version__
from ..common.context import AppContext
from ..common.globals import APPNAME
from . import globals as defaults
from .configuration_manager import NodeConfigurationManager, ServerConfigurationManager

_UNSET = object()


def split_db_uri(uri: str):
    """``(is_file_based, path_or_none)`` for a SQLAlchemy-style URI (``sqlite:///rel.db``,
    ``sqlite:////abs/path.db``, ``postgresql://host/db``) without needing SQLAlchemy."""
    parts = urlsplit(uri)
    if parts.scheme.startswith("sqlite"):
        return True, (uri.split(":///", 1)[1] if ":///" in uri else "")
    return (parts.hostname is None), (parts.path or None)


class _BoundContext(AppContext):
    """An ``AppContext`` whose instance type and defaults are class attributes."""

    INSTANCE_TYPE = ""
    DEFAULT_ENVIRONMENT = ""
    DEFAULT_SYSTEM_FOLDERS = False

    def __init__(self, instance_name, environment=_UNSET, system_folders=_UNSET, config_file=None):
        env, folders = self._defaults(environment, system_folders)
        super().__init__(self.INSTANCE_TYPE, instance_name, env, folders, config_file)
        self.log.info(f"vantage6 version '{__version__}'")

    @classmethod
    def _defaults(cls, environment, system_folders):
        return (cls.DEFAULT_ENVIRONMENT if environment is _UNSET else environment,
                cls.DEFAULT_SYSTEM_FOLDERS if system_folders is _UNSET else system_folders)

    @classmethod
    def from_external_config_file(cls, path, environment=_UNSET, system_folders=_UNSET):
        env, folders = cls._defaults(environment, system_folders)
        return super().from_external_config_file(path, cls.INSTANCE_TYPE, env, folders)

    @classmethod
    def config_exists(cls, instance_name, environment=_UNSET, system_folders=_UNSET):
        env, folders = cls._defaults(environment, system_folders)
        return super().config_exists(cls.INSTANCE_TYPE, instance_name, environment=env, system_folders=folders)

    @classmethod
    def available_configurations(cls, system_folders=_UNSET):
        _, folders = cls._defaults(_UNSET, system_folders)
        return super().available_configurations(cls.INSTANCE_TYPE, folders)

    @property
    def _runtime_base(self) -> str:
        return f"{APPNAME}-{self.name}-{self.scope}"


class ServerContext(_BoundContext):
    """Context of a central server; knows where its database lives."""

    INSTANCE_TYPE = "server"
    DEFAULT_ENVIRONMENT = defaults.DEFAULT_SERVER_ENVIRONMENT
    DEFAULT_SYSTEM_FOLDERS = defaults.DEFAULT_SERVER_SYSTEM_FOLDERS
    INST_CONFIG_MANAGER = ServerConfigurationManager

    def __init__(self, instance_name, environment=_UNSET, system_folders=_UNSET):
        super().__init__(instance_name, environment, system_folders)

    def get_database_uri(self):
        """``VANTAGE6_DB_URI`` wins over ``config['uri']``; a relative sqlite path is anchored at ``data_dir``."""
        uri = os.environ.get("VANTAGE6_DB_URI") or self.config["uri"]
        is_file, path = split_db_uri(uri)
        if is_file and path and not os.path.isabs(path):
            uri = f"{uri.split(':///', 1)[0]}:///{self.data_dir / path}"
        return uri

    @property
    def docker_container_name(self):
        return f"{self._runtime_base}-server"

    @classmethod
    def from_external_config_file(cls, path, environment=_UNSET, system_folders=_UNSET):
        ctx = super().from_external_config_file(path, environment, system_folders)
        ctx.name = os.environ.get("VANTAGE6_CONFIG_NAME") or ctx.name      # set by `vserver start` for the runtime
        return ctx


class NodeContext(_BoundContext):
    """Context of a node, used by the CLI on the host and by the node runtime itself."""

    INSTANCE_TYPE = "node"
    DEFAULT_ENVIRONMENT = defaults.DEFAULT_NODE_ENVIRONMENT
    DEFAULT_SYSTEM_FOLDERS = defaults.DEFAULT_NODE_SYSTEM_FOLDERS
    INST_CONFIG_MANAGER = NodeConfigurationManager
    running_in_docker = False

    @staticmethod
    def type_data_folder(system_folders):
        return AppContext.type_data_folder("node", system_folders)

    @property
    def databases(self):
        return self.config["databases"]

    def get_database_uri(self, label="default"):
        return self.databases[label]

    # names of the runtime objects that belong to this node
    @property
    def docker_container_name(self):
        return self._runtime_base

    @property
    def docker_network_name(self):
        return f"{self._runtime_base}-net"

    @property
    def docker_volume_name(self):
        return os.environ.get("DATA_VOLUME_NAME", f"{self._runtime_base}-vol")

    @property
    def docker_vpn_volume_name(self):
        return os.environ.get("VPN_VOLUME_NAME", f"{self._runtime_base}-vpn-vol")

    def docker_temporary_volume_name(self, run_id):
        return f"{self._runtime_base}-{run_id}-tmpvol"
