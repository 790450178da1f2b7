"""Helpers around the process runtime -- the ``vantage6.common.docker.addons`` /
``network_manager`` contract used by the reference CLI (reference vantage6/cli/node.py:38-42,
vantage6/cli/server.py:16-20, rabbitmq/queue_manager.py:14-15)."""
from __future__ import annotations

import json
import re

from . import APIError, LocalRuntime, NotFound, from_env, runtime_dir
from ..common import error
from ..common.globals import APPNAME


def check_docker_running() -> None:
    """Exit(1) with an error when the runtime is unusable (name kept for CLI parity:
    reference vantage6/cli/node.py:76)."""
    try:
        from_env().ping()
    except Exception as e:  # noqa: BLE001
        error("Cannot reach the process runtime! Please make sure the runtime directory is writable.")
        error(str(e))
        exit(1)


check_runtime_running = check_docker_running


def pull_if_newer(runtime: LocalRuntime, image: str, log=None) -> None:
    """Images are entry points shipped with this package: nothing to download. Unknown,
    non-reference image names raise so the caller prints its "alas, no dice" warning."""
    if not image or not re.match(r"^[\w./:@-]+$", image):
        raise APIError(f"invalid image reference {image!r}")
    runtime.images.pull(image)


def remove_container_if_exists(docker_client: LocalRuntime, **filters) -> None:
    name = filters.get("name")
    if name is None:
        return
    try:
        c = docker_client.containers.get(name)
    except NotFound:
        return
    try:
        c.remove(force=True)
    except Exception:  # noqa: BLE001
        pass


def remove_container(container, kill: bool = False) -> None:
    try:
        container.remove(force=kill)
    except Exception as e:  # noqa: BLE001
        error(f"Failed to remove container {container.name}: {e}")


def get_server_config_name(container_name: str, scope: str) -> str:
    """``vantage6-{name}-{scope}-server`` -> ``{name}`` (reference server.py:605-608)."""
    idx_scope = container_name.rfind(scope)
    length_app_name = len(APPNAME)
    return container_name[length_app_name + 1: idx_scope - 1]


class NetworkManager:
    """Named "network" = a registry entry other services can join; for processes on one box the
    network is the loopback interface, so this only records membership
    (reference vantage6/cli/server.py:210-213: ``NetworkManager(network_name).create_network(is_internal)``)."""

    def __init__(self, network_name: str):
        self.network_name = network_name
        self._file = runtime_dir() / "networks" / f"{network_name}.json"

    def create_network(self, is_internal: bool = True) -> None:
        self._file.parent.mkdir(parents=True, exist_ok=True)
        if not self._file.exists():
            self._file.write_text(json.dumps({"name": self.network_name, "internal": is_internal, "members": []}))

    def delete_network(self, kill_containers: bool = True) -> None:
        try:
            self._file.unlink()
        except FileNotFoundError:
            pass

    def connect(self, container_name: str, aliases=None, ipv4=None) -> None:
        self.create_network()
        data = json.loads(self._file.read_text())
        if container_name not in data["members"]:
            data["members"].append(container_name)
        self._file.write_text(json.dumps(data))

    def contains(self, container) -> bool:
        if not self._file.exists():
            return False
        return getattr(container, "name", container) in json.loads(self._file.read_text())["members"]
