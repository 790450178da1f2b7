"""Defaults of the command line tools, one record per instance type.

The flat ``DEFAULT_*`` names are what the rest of the package (and code written against vantage6)
imports; values and names follow the reference (vantage6/cli/globals.py:10-35, including its
``PACAKAGE_FOLDER`` spelling).
"""
from pathlib import Path
from typing import NamedTuple

from ..common.globals import APPNAME, STRING_ENCODING  # noqa: F401  (re-exported)


class InstanceDefaults(NamedTuple):
    environment: str           # configuration environment used when -e is not given
    system_folders: bool       # True: system-wide folders, False: the user's folders


SERVER_DEFAULTS = InstanceDefaults(environment="prod", system_folders=True)
NODE_DEFAULTS = InstanceDefaults(environment="application", system_folders=False)

DEFAULT_SERVER_ENVIRONMENT, DEFAULT_SERVER_SYSTEM_FOLDERS = SERVER_DEFAULTS
DEFAULT_NODE_ENVIRONMENT, DEFAULT_NODE_SYSTEM_FOLDERS = NODE_DEFAULTS

# where the package lives / keeps bundled data
PACAKAGE_FOLDER = Path(__file__).resolve().parents[2]
DATA_FOLDER = PACAKAGE_FOLDER / APPNAME / "_data"

NODE_PROXY_SERVER_HOSTNAME = "proxyserver"     # host name algorithms use to reach their node's proxy
RABBIT_TIMEOUT = 5 * 60                        # seconds the message-queue sidecar gets to come up
