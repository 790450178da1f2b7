"""Global constants (the ``vantage6.common.globals`` contract seen from the reference CLI:
reference vantage6/cli/node.py:31-37, vantage6/cli/server.py:21-26, vantage6/cli/globals.py:2-5)."""
from pathlib import Path

APPNAME = "vantage6"
STRING_ENCODING = "utf-8"
PACKAGE_FOLDER = Path(__file__).parent.parent.parent

# "images" are python entry points in this framework; the registry/name scheme is kept so
# that reference configuration files (which may carry an `image:` key) stay valid.
DEFAULT_DOCKER_REGISTRY = "harbor2.vantage6.ai"
DEFAULT_NODE_IMAGE = "infrastructure/node:petronas"
DEFAULT_SERVER_IMAGE = "infrastructure/server:petronas"

VPN_CONFIG_FILE = "vpn-config.ovpn.conf"
DATABASE_TYPES = ["csv", "parquet", "sql", "sparql", "excel", "other", "pt", "npy"]

# environment variable that relocates every config/data/log folder (tests, sandboxes)
HOME_ENV = "V6B200_HOME"
