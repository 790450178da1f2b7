"""Server / node contexts (reference vantage6/cli/context.py:16-144): names of the running
instance ("container"), its network and volumes, and the database URI resolution with the
``VANTAGE6_DB_URI`` / ``VANTAGE6_CONFIG_NAME`` / ``DATA_VOLUME_NAME`` / ``VPN_VOLUME_NAME``
environment overrides."""
import os.path
from urllib.parse import urlsplit

from .._version import __version__
from ..common.context import AppContext
from ..common.globals import APPNAME
from .configuration_manager import NodeConfigurationManager, ServerConfigurationManager
from .globals import DEFAULT_NODE_ENVIRONMENT as N_ENV
from .globals import DEFAULT_NODE_SYSTEM_FOLDERS as N_FOL
from .globals import DEFAULT_SERVER_ENVIRONMENT as S_ENV
from .globals import DEFAULT_SERVER_SYSTEM_FOLDERS as S_FOL


def split_db_uri(uri: str):
    """``(is_file_based, path_or_none)`` for a SQLAlchemy-style URI (``sqlite:///rel.db``,
    ``sqlite:////abs/path.db``, ``postgresql://host/db``) without needing SQLAlchemy."""
    parts = urlsplit(uri)
    if parts.scheme.startswith("sqlite"):
        path = uri.split(":///", 1)[1] if ":///" in uri else ""
        return True, path
    return (parts.hostname is None), (parts.path or None)


class ServerContext(AppContext):
    """Context for the server: overrides where the database lives."""

    INST_CONFIG_MANAGER = ServerConfigurationManager

    def __init__(self, instance_name, environment=S_ENV, system_folders=S_FOL):
        super().__init__("server", instance_name, environment=environment, system_folders=system_folders)
        self.log.info(f"vantage6 version '{__version__}'")

    def get_database_uri(self):
        """``VANTAGE6_DB_URI`` overrides ``config['uri']``; relative sqlite paths resolve
        against ``data_dir`` (reference context.py:30-42)."""
        uri = os.environ.get("VANTAGE6_DB_URI") or self.config["uri"]
        file_based, path = split_db_uri(uri)
        if file_based and path and not os.path.isabs(path):
            scheme = uri.split(":///", 1)[0]
            uri = f"{scheme}:///{self.data_dir / path}"
        return uri

    @property
    def docker_container_name(self):
        return f"{APPNAME}-{self.name}-{self.scope}-server"

    @classmethod
    def from_external_config_file(cls, path, environment=S_ENV, system_folders=S_FOL):
        cls_ = super().from_external_config_file(path, "server", environment, system_folders)
        # a server started by the runtime gets its config name from the environment
        cls_.name = os.environ.get("VANTAGE6_CONFIG_NAME") or cls_.name
        return cls_

    @classmethod
    def config_exists(cls, instance_name, environment=S_ENV, system_folders=S_FOL):
        return super().config_exists("server", instance_name, environment=environment, system_folders=system_folders)

    @classmethod
    def available_configurations(cls, system_folders=S_FOL):
        return super().available_configurations("server", system_folders)


class NodeContext(AppContext):
    """Node context on the host (used by the CLI and by the node runtime)."""

    INST_CONFIG_MANAGER = NodeConfigurationManager
    running_in_docker = False

    def __init__(self, instance_name, environment=N_ENV, system_folders=N_FOL, config_file=None):
        super().__init__("node", instance_name, environment, system_folders, config_file)
        self.log.info(f"vantage6 version '{__version__}'")

    @classmethod
    def from_external_config_file(cls, path, environment=N_ENV, system_folders=N_FOL):
        return super().from_external_config_file(path, "node", environment, system_folders)

    @classmethod
    def config_exists(cls, instance_name, environment=N_ENV, system_folders=N_FOL):
        return super().config_exists("node", instance_name, environment=environment, system_folders=system_folders)

    @classmethod
    def available_configurations(cls, system_folders=N_FOL):
        return super().available_configurations("node", system_folders)

    @staticmethod
    def type_data_folder(system_folders):
        return AppContext.type_data_folder("node", system_folders)

    @property
    def databases(self):
        return self.config["databases"]

    @property
    def docker_container_name(self):
        return f"{APPNAME}-{self.name}-{self.scope}"

    @property
    def docker_network_name(self):
        return f"{APPNAME}-{self.name}-{self.scope}-net"

    @property
    def docker_volume_name(self):
        return os.environ.get("DATA_VOLUME_NAME", f"{self.docker_container_name}-vol")

    @property
    def docker_vpn_volume_name(self):
        return os.environ.get("VPN_VOLUME_NAME", f"{self.docker_container_name}-vpn-vol")

    def docker_temporary_volume_name(self, run_id):
        return f"{APPNAME}-{self.name}-{self.scope}-{run_id}-tmpvol"

    def get_database_uri(self, label="default"):
        return self.config["databases"][label]
