"""CLI helpers (reference vantage6/cli/utils.py:6-19)."""
import re

from ..common import error


def check_config_name_allowed(name: str) -> None:
    """Configuration names double as process / volume names: ``^[a-zA-Z0-9_.-]+$`` only."""
    if not re.match("^[a-zA-Z0-9_.-]+$", name):
        error(f"Name '{name}' is not allowed. Please use only the following characters: a-zA-Z0-9_.-")
        exit(1)


def check_if_docker_deamon_is_running(docker_client) -> None:
    """Name kept for parity; checks the process runtime (reference utils.py:14-19)."""
    try:
        docker_client.ping()
    except Exception:  # noqa: BLE001
        error("Docker socket can not be found. Make sure Docker is running.")
        exit(1)
