"""The message-queue sidecar of a server.

Why it exists (reference vantage6/cli/server.py:267-268): without a message queue "the server application
cannot scale horizontally" -- several server processes must see each other's events.  The reference runs a
RabbitMQ container next to the server; here the sidecar is the in-box ZeroMQ forwarder
(``v6-mq-broker``, server/mq_broker.py) run by the process runtime, with the lifecycle the reference gives
its container (vantage6/cli/rabbitmq/queue_manager.py:25-238):

* ``amqp://user:password@host:port/vhost`` from the ``rabbitmq_uri`` configuration key says who / where;
* ``definitions.json`` (one administrator with a salted-SHA256 password hash, one vhost, full permissions) and
  ``rabbitmq.config`` are written into the server's data directory and handed to the sidecar read-only, next to
  a persistent ``<data>/rabbitmq`` directory;
* the sidecar is called ``vantage6-{name}-rabbitmq``, labelled ``vantage6-type=rabbitmq``, restarted always, and
  gets ``RABBIT_TIMEOUT`` seconds (polled every ``INTERVAL``) to answer a status probe before the CLI gives up
  with exit code 1.
"""
from __future__ import annotations

import base64
import copy
import hashlib
import json
import os
import shutil
import time
from dataclasses import dataclass
from pathlib import Path
from typing import Dict

from ... import runtime as docker
from ...common import debug, error, info
from ...common.globals import APPNAME
from ...runtime.addons import NetworkManager, remove_container_if_exists
from ..globals import RABBIT_TIMEOUT
from .definitions import RABBITMQ_DEFINITIONS

DEFAULT_RABBIT_IMAGE = "harbor2.vantage6.ai/infrastructure/rabbitmq"
RABBIT_CONFIG = "rabbitmq.config"
RABBIT_DIR = "rabbitmq"
_SALT_BYTES = 4                                 # RabbitMQ salts its password hashes with 32 bits
_INSIDE = {"definitions": "/etc/rabbitmq/definitions.json", "config": f"/etc/rabbitmq/{RABBIT_CONFIG}",
           "data": "/var/lib/rabbitmq"}


@dataclass(frozen=True)
class QueueAddress:
    user: str
    password: str
    host: str
    port: str
    vhost: str

    @classmethod
    def parse(cls, uri: str) -> "QueueAddress":
        """``amqp://$user:$pass@$host:$port/$vhost``; the password may itself contain ``:``."""
        credentials, _, location = uri.partition("@")
        user, _, password = credentials.rsplit("/", 1)[-1].partition(":")
        host, _, rest = location.partition(":")
        port, _, vhost = rest.partition("/")
        return cls(user, password, host, port, vhost)


def split_rabbitmq_uri(rabbit_uri: str) -> Dict[str, str]:
    """The parts of a queue URI as a plain dict (``user password host port vhost``)."""
    return dict(QueueAddress.parse(rabbit_uri).__dict__)


class RabbitMQManager:
    """Writes the sidecar's files, starts it on the server's network and waits until it answers."""

    INTERVAL = 10           # seconds between status probes while waiting for the sidecar

    def __init__(self, ctx, network_mgr: NetworkManager, image: str = None) -> None:
        self.ctx = ctx
        self.network_mgr = network_mgr
        self.queue_uri = ctx.config.get("rabbitmq_uri")
        address = QueueAddress.parse(self.queue_uri)
        self.rabbit_user, self.rabbit_pass = address.user, address.password
        self.host, self.port, self.vhost = address.host, address.port, address.vhost
        self.image = image or DEFAULT_RABBIT_IMAGE
        self.docker = docker.from_env()
        self.rabbit_container_name = f"{APPNAME}-{ctx.name}-rabbitmq"
        self.definitions_file = Path(ctx.data_dir / "definitions.json")

    # -------------------------------------------------------------------------------- lifecycle
    def start(self) -> None:
        volumes = self._get_volumes()
        remove_container_if_exists(docker_client=self.docker, name=self.rabbit_container_name)     # stale sidecar
        command = (f"v6-mq-broker serve --port {self.port} --definitions {_INSIDE['definitions']} "
                   f"--data {_INSIDE['data']}")
        self.rabbit_container = self.docker.containers.run(
            image=self.image, name=self.rabbit_container_name, hostname=self.rabbit_container_name, command=command,
            volumes=volumes, ports={f"{self.port}/tcp": self.port, "15672/tcp": 8080}, detach=True, auto_remove=False,
            restart_policy={"Name": "always"}, labels={f"{APPNAME}-type": "rabbitmq"},
            network=self.network_mgr.network_name)
        self._wait_for_startup()

    def is_running(self) -> bool:
        probe = self.rabbit_container.exec_run(cmd=f"rabbitmqctl status --formatter json --port {self.port}")
        return probe.exit_code == 0

    def _wait_for_startup(self) -> None:
        deadline_probes = int((RABBIT_TIMEOUT + self.INTERVAL) / self.INTERVAL)
        for _ in range(deadline_probes):
            if self.is_running():
                info("RabbitMQ was started successfully!")
                return
            debug(f"RabbitMQ is not yet running. Retrying in {self.INTERVAL}s...")
            time.sleep(self.INTERVAL)
        error("Could not start RabbitMQ! Exiting...")
        exit(1)

    # ---------------------------------------------------------------------------- files on disk
    def _get_volumes(self) -> Dict:
        """Write definitions + config into the data directory; returns the bind specification."""
        data_dir = Path(self.ctx.data_dir)
        data_dir.mkdir(parents=True, exist_ok=True)
        self.definitions_file.write_text(json.dumps(self._get_rabbitmq_definitions(), indent=2))
        config_copy = data_dir / RABBIT_CONFIG
        shutil.copyfile(Path(__file__).resolve().with_name(RABBIT_CONFIG), config_copy)
        state_dir = data_dir / RABBIT_DIR
        state_dir.mkdir(parents=True, exist_ok=True)
        return {self.definitions_file: {"bind": _INSIDE["definitions"], "mode": "ro"},
                config_copy: {"bind": _INSIDE["config"], "mode": "ro"},
                state_dir: {"bind": _INSIDE["data"], "mode": "rw"}}

    def _get_rabbitmq_definitions(self) -> Dict:
        """The definitions template with this queue's user, password hash and vhost filled in."""
        doc = copy.deepcopy(RABBITMQ_DEFINITIONS)
        (admin,), (vhost,), (grant,) = doc["users"], doc["vhosts"], doc["permissions"]
        admin.update(name=self.rabbit_user, password_hash=self._get_hashed_pw(self.rabbit_pass))
        vhost.update(name=self.vhost)
        grant.update(user=self.rabbit_user, vhost=self.vhost)
        return doc

    # -------------------------------------------------------------------------- password hashing
    @staticmethod
    def _digest(salt: bytes, pw: str) -> bytes:
        return hashlib.sha256(salt + pw.encode("utf-8")).digest()

    @classmethod
    def _get_hashed_pw(cls, pw: str) -> str:
        """RabbitMQ's ``rabbit_password_hashing_sha256``: base64(salt || sha256(salt || utf8(pw)))."""
        salt = os.urandom(_SALT_BYTES)
        return base64.b64encode(salt + cls._digest(salt, pw)).decode("utf-8")

    @classmethod
    def check_pw(cls, pw: str, hashed: str) -> bool:
        raw = base64.b64decode(hashed)
        return cls._digest(raw[:_SALT_BYTES], pw) == raw[_SALT_BYTES:]
