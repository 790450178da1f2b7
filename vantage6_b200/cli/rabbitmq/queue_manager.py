"""Message-queue sidecar manager (behavioural spec: reference
vantage6/cli/rabbitmq/queue_manager.py:25-238).

Purpose (reference vantage6/cli/server.py:267-268): without a message queue "the server
application cannot scale horizontally" -- several server processes must share events.  The
reference starts a RabbitMQ container; here the sidecar is an in-box ZeroMQ forwarder process
(server/mq_broker.py) managed by the process runtime with the same lifecycle: definitions +
config written into the server data dir, persistent dir ``<data>/rabbitmq``, label
``vantage6-type=rabbitmq``, name ``vantage6-{name}-rabbitmq``, restart policy "always", start-up
poll every 10 s up to ``RABBIT_TIMEOUT`` (300 s) then ``exit(1)``.
"""
from __future__ import annotations

import base64
import copy
import hashlib
import json
import os
import shutil
import time
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


def split_rabbitmq_uri(rabbit_uri: str) -> Dict[str, str]:
    """``amqp://$user:$pass@$host:$port/$vhost`` -> its parts."""
    user_details, location_details = rabbit_uri.split("@", 1)
    user, password = user_details.split("/")[-1].split(":", 1)
    host, remainder = location_details.split(":", 1)
    port, vhost = remainder.split("/", 1)
    return {"user": user, "password": password, "host": host, "port": port, "vhost": vhost}


class RabbitMQManager:
    """Manages the message-queue sidecar process."""

    INTERVAL = 10

    def __init__(self, ctx, network_mgr: NetworkManager, image: str = None) -> None:
        self.ctx = ctx
        self.queue_uri = self.ctx.config.get("rabbitmq_uri")
        parts = split_rabbitmq_uri(self.queue_uri)
        self.rabbit_user, self.rabbit_pass = parts["user"], parts["password"]
        self.vhost, self.port, self.host = parts["vhost"], parts["port"], parts["host"]
        self.definitions_file = Path(self.ctx.data_dir / "definitions.json")
        self.network_mgr = network_mgr
        self.docker = docker.from_env()
        self.image = image if image else DEFAULT_RABBIT_IMAGE
        self.rabbit_container_name = f"{APPNAME}-{ctx.name}-rabbitmq"

    def start(self) -> None:
        volumes = self._get_volumes()
        ports = {f"{self.port}/tcp": self.port, "15672/tcp": 8080}
        # a sidecar left over from a previous run is replaced
        remove_container_if_exists(docker_client=self.docker, name=self.rabbit_container_name)
        self.rabbit_container = self.docker.containers.run(
            name=self.rabbit_container_name, image=self.image,
            command=f"v6-mq-broker serve --port {self.port} --definitions /etc/rabbitmq/definitions.json "
                    f"--data /var/lib/rabbitmq",
            volumes=volumes, ports=ports, detach=True, restart_policy={"Name": "always"},
            hostname=f"{APPNAME}-{self.ctx.name}-rabbitmq", labels={f"{APPNAME}-type": "rabbitmq"},
            network=self.network_mgr.network_name, auto_remove=False)
        self._wait_for_startup()

    def _wait_for_startup(self) -> None:
        attempts = int((RABBIT_TIMEOUT + self.INTERVAL) / self.INTERVAL)
        for _ in range(attempts):
            if self.is_running():
                info("RabbitMQ was started successfully!")
                return
            debug(f"RabbitMQ is not yet running. Retrying in {self.INTERVAL}s...")
            time.sleep(self.INTERVAL)
        error("Could not start RabbitMQ! Exiting...")
        exit(1)

    def is_running(self) -> bool:
        response = self.rabbit_container.exec_run(cmd=f"rabbitmqctl status --formatter json --port {self.port}")
        return response.exit_code == 0

    def _get_volumes(self) -> Dict:
        Path(self.ctx.data_dir).mkdir(parents=True, exist_ok=True)
        with open(self.definitions_file, "w") as f:
            json.dump(self._get_rabbitmq_definitions(), f, indent=2)
        shutil.copyfile(Path(__file__).parent.resolve() / RABBIT_CONFIG, self.ctx.data_dir / RABBIT_CONFIG)
        rabbit_data_dir = self.ctx.data_dir / RABBIT_DIR
        rabbit_data_dir.mkdir(parents=True, exist_ok=True)
        return {
            self.definitions_file: {"bind": "/etc/rabbitmq/definitions.json", "mode": "ro"},
            self.ctx.data_dir / RABBIT_CONFIG: {"bind": "/etc/rabbitmq/rabbitmq.config", "mode": "ro"},
            rabbit_data_dir: {"bind": "/var/lib/rabbitmq", "mode": "rw"},
        }

    def _get_rabbitmq_definitions(self) -> Dict:
        d = copy.deepcopy(RABBITMQ_DEFINITIONS)
        d["users"][0]["name"] = self.rabbit_user
        d["permissions"][0]["user"] = self.rabbit_user
        d["users"][0]["password_hash"] = self._get_hashed_pw(self.rabbit_pass)
        d["vhosts"][0]["name"] = self.vhost
        d["permissions"][0]["vhost"] = self.vhost
        return d

    @staticmethod
    def _get_hashed_pw(pw: str) -> str:
        """RabbitMQ's salted SHA-256: b64(salt || sha256(salt || utf8(pw))) with a 32-bit salt."""
        salt = os.urandom(4)
        digest = hashlib.sha256(salt + pw.encode("utf-8")).digest()
        return base64.b64encode(salt + digest).decode("utf-8")

    @staticmethod
    def check_pw(pw: str, hashed: str) -> bool:
        raw = base64.b64decode(hashed)
        return hashlib.sha256(raw[:4] + pw.encode("utf-8")).digest() == raw[4:]
