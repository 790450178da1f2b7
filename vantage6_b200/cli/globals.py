"""CLI defaults (reference vantage6/cli/globals.py:10-35)."""
from pathlib import Path

from ..common.globals import APPNAME, STRING_ENCODING  # noqa: F401

# server settings
DEFAULT_SERVER_SYSTEM_FOLDERS = True
DEFAULT_SERVER_ENVIRONMENT = "prod"

# node settings
DEFAULT_NODE_SYSTEM_FOLDERS = False
DEFAULT_NODE_ENVIRONMENT = "application"

# installation settings
PACAKAGE_FOLDER = Path(__file__).parent.parent.parent
NODE_PROXY_SERVER_HOSTNAME = "proxyserver"
DATA_FOLDER = PACAKAGE_FOLDER / APPNAME / "_data"

# maximum time to start up the message-queue sidecar, in seconds
RABBIT_TIMEOUT = 300
