"""Demo network: one server + N nodes on this box, driven through the real CLI commands
(``vserver import/start``, ``vnode start``) and the process runtime.

Used by the CPU plumbing tests (BASELINE config 1: 2 CPU nodes + 1 server process) and on the
GPU box (8 nodes, node k pinned to GPU k).  Everything lives under ``home`` (``$V6B200_HOME``).

    net = DemoNetwork(n_nodes=2, home=tmpdir).start()
    client = net.client()
    task = client.task.create(collaboration=net.collaboration_id, organizations=[net.org_ids[0]],
                              name="mean", image="v6b200/weighted-mean",
                              input={"method": "master", "master": True})
    results = client.wait_for_results(task["id"])
    net.stop()
"""
from __future__ import annotations

import os
import socket
import time
import uuid
from pathlib import Path
from typing import List, Optional

import yaml
from click.testing import CliRunner

from .common.globals import HOME_ENV

LOGGING = {"level": "INFO", "file": "x.log", "use_console": True, "backup_count": 5, "max_size": 1024,
           "format": "%(asctime)s - %(name)-14s - %(levelname)-8s - %(message)s", "datefmt": "%Y-%m-%d %H:%M:%S"}


def free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class DemoNetwork:
    def __init__(self, n_nodes: int = 2, home: Optional[str] = None, name: str = "demo", gpus: Optional[List[int]] = None,
                 databases: Optional[List[str]] = None, encrypted: bool = False, rabbitmq: bool = False, tls: bool = False):
        self.n_nodes, self.name = n_nodes, name
        self.home = Path(home or os.environ.get(HOME_ENV) or Path.cwd() / ".v6b200")
        os.environ[HOME_ENV] = str(self.home)
        self.gpus = gpus
        self.databases = databases or [""] * n_nodes
        self.encrypted = encrypted
        self.rabbitmq = rabbitmq
        self.tls = tls                           # https / wss with a self-signed certificate the nodes and clients verify against
        self.port = free_port()
        self.api_keys = [str(uuid.uuid4()) for _ in range(n_nodes)]
        self.org_names = [f"org-{i}" for i in range(n_nodes)]
        self.org_ids: List[int] = []
        self.collaboration_id: Optional[int] = None
        self.password = "demo-password"
        self.key_bits = 2048                     # demo keys; `vnode create-private-key` makes 4096-bit ones
        self.log: List[str] = []

    # ------------------------------------------------------------------ config files
    def _cfg_dir(self, kind: str) -> Path:
        d = self.home / "user" / "config" / kind
        d.mkdir(parents=True, exist_ok=True)
        return d

    @property
    def server_url(self) -> str:
        return ("https" if self.tls else "http") + "://127.0.0.1"

    @property
    def cert_file(self) -> str:
        return str(self.home / "keys" / "server.crt")

    def key_file(self, org: int) -> str:
        """RSA key of organization ``org`` (shared by its node and its researchers, as in vantage6); created on first use."""
        path = self.home / "keys" / f"privkey_{self.org_names[org]}.pem"
        if not path.exists():
            from .common.encryption import RSACryptor

            RSACryptor.create_new_rsa_key(path, bits=self.key_bits)
        return str(path)

    def write_configs(self) -> None:
        server_cfg = {"description": "demo network", "ip": "127.0.0.1", "port": self.port, "api_path": "/api",
                      "uri": "sqlite:///demo.sqlite", "allow_drop_all": True, "jwt_secret_key": str(uuid.uuid4()),
                      "logging": dict(LOGGING, file=f"{self.name}.log")}
        if self.tls:
            from .common.encryption import create_self_signed_certificate

            if not Path(self.cert_file).exists():
                create_self_signed_certificate(self.cert_file, str(self.home / "keys" / "server.key"))
            server_cfg["ssl"] = {"certfile": self.cert_file, "keyfile": str(self.home / "keys" / "server.key")}
        if self.rabbitmq:
            server_cfg["rabbitmq_uri"] = f"amqp://demo:demo@127.0.0.1:{free_port()}/demo"
        with open(self._cfg_dir("server") / f"{self.name}.yaml", "w") as f:
            yaml.safe_dump({"application": {}, "environments": {"prod": server_cfg, "acc": {}, "test": {}, "dev": {}}}, f)
        for i in range(self.n_nodes):
            node_cfg = {"api_key": self.api_keys[i], "server_url": self.server_url, "port": self.port, "api_path": "/api",
                        **({"server_ca_file": self.cert_file} if self.tls else {}),
                        "task_dir": str(self.home / "tasks" / f"node-{i}"),
                        "databases": (dict(self.databases[i]) if isinstance(self.databases[i], dict)      # {label: uri}
                                      else {"default": self.databases[i] or f"synthetic://node-{i}"}),
                        "logging": dict(LOGGING, file=f"node-{i}.log"),
                        "encryption": {"enabled": self.encrypted, "private_key": self.key_file(i) if self.encrypted else ""}}
            if self.gpus is not None:
                node_cfg["gpu"] = self.gpus[i]
            with open(self._cfg_dir("node") / f"{self.name}-node-{i}.yaml", "w") as f:
                yaml.safe_dump({"application": node_cfg, "environments": {"prod": {}, "acc": {}, "test": {}, "dev": {}}}, f)
        fixtures = {
            "organizations": [{"name": n, "domain": f"{n}.test",
                               "users": [{"username": f"user-{i}", "password": self.password, "firstname": "u",
                                          "lastname": str(i), "roles": ["Root"] if i == 0 else ["Researcher"]}]}
                              for i, n in enumerate(self.org_names)],
            "collaborations": [{"name": f"{self.name}-collab", "encrypted": self.encrypted,
                                "participants": [{"name": n, "api-key": k} for n, k in zip(self.org_names, self.api_keys)]}],
        }
        self.fixtures_file = self.home / f"{self.name}-entities.yaml"
        with open(self.fixtures_file, "w") as f:
            yaml.safe_dump(fixtures, f)

    # ------------------------------------------------------------------ lifecycle
    def _invoke(self, cmd, args):
        r = CliRunner().invoke(cmd, args, catch_exceptions=False)
        self.log.append(r.output)
        if r.exit_code != 0:
            raise RuntimeError(f"{cmd.name} {' '.join(args)} failed ({r.exit_code}):\n{r.output}")
        return r

    # -- persistence: `vdev create-demo-network` and `vdev start-demo-network` are separate invocations --------------
    def description_file(self) -> Path:
        return self.home / f"{self.name}-network.json"

    def save(self) -> Path:
        import json

        doc = {k: getattr(self, k) for k in ("n_nodes", "name", "gpus", "databases", "encrypted", "rabbitmq", "tls", "port", "api_keys",
                                             "org_names", "password", "key_bits")}
        self.description_file().write_text(json.dumps(doc, indent=1))
        return self.description_file()

    @classmethod
    def load(cls, name: str = "demo", home: Optional[str] = None) -> "DemoNetwork":
        import json

        root = Path(home or os.environ.get(HOME_ENV) or Path.cwd() / ".v6b200")
        doc = json.loads((root / f"{name}-network.json").read_text())
        net = cls(doc["n_nodes"], home=str(root), name=doc["name"], gpus=doc["gpus"], databases=doc["databases"],
                  encrypted=doc["encrypted"], rabbitmq=doc["rabbitmq"], tls=doc.get("tls", False))
        for k in ("port", "api_keys", "org_names", "password", "key_bits"):
            setattr(net, k, doc[k])
        net.fixtures_file = net.home / f"{net.name}-entities.yaml"
        return net

    def create(self, timeout: float = 60.0) -> "DemoNetwork":
        """Configuration files, keys, the entities file and the imported database -- everything but running processes."""
        from .cli.server import cli_server_import
        from .runtime import from_env

        self.write_configs()
        self.save()
        self._invoke(cli_server_import, ["--user", "-n", self.name, "--drop-all", str(self.fixtures_file)])
        rt = from_env()
        t0 = time.time()
        while any(c.labels.get("name") == self.name and "import" in " ".join(c.meta["command"])
                  for c in rt.containers.list(filters={"label": "vantage6-type=server"})):
            if time.time() - t0 > timeout:
                raise TimeoutError("vserver import did not finish")
            time.sleep(0.1)
        return self

    def start(self, timeout: float = 60.0) -> "DemoNetwork":
        return self.create(timeout).up(timeout)

    def up(self, timeout: float = 60.0) -> "DemoNetwork":
        """Start the server and the nodes of a created network."""
        from .cli.node import cli_node_start
        from .cli.server import cli_server_start

        self._invoke(cli_server_start, ["--user", "-n", self.name])
        client = self.client(timeout=timeout)
        orgs = {o["name"]: o["id"] for o in client.organization.list()}
        self.org_ids = [orgs[n] for n in self.org_names]
        self.collaboration_id = next(c["id"] for c in client.collaboration.list() if c["name"] == f"{self.name}-collab")
        for i in range(self.n_nodes):
            args = ["--user", "-n", f"{self.name}-node-{i}"]
            if self.gpus is not None:
                args += ["--gpu", str(self.gpus[i])]
            self._invoke(cli_node_start, args)
        t0 = time.time()
        while True:
            online = [n for n in client.node.list() if n["status"] == "online"]
            # a node is online as soon as it has a token; in an encrypted collaboration it is usable once it has also
            # published its organization's public key (the next thing it does)
            keyed = not self.encrypted or all(o.get("public_key") for o in client.organization.list() if o["name"] in self.org_names)
            if len(online) >= self.n_nodes and keyed:
                break
            if time.time() - t0 > timeout:
                raise TimeoutError(f"only {len(online)}/{self.n_nodes} nodes came online:\n{self.tail_logs()}")
            time.sleep(0.2)
        return self

    def client(self, user: int = 0, timeout: float = 30.0):
        from .client import UserClient

        c = UserClient(self.server_url, self.port, "/api", ca_file=self.cert_file if self.tls else None)
        t0 = time.time()
        while True:
            try:
                c.authenticate(f"user-{user}", self.password)
                break
            except Exception:  # noqa: BLE001 -- server still starting
                if time.time() - t0 > timeout:
                    raise
                time.sleep(0.2)
        c.setup_encryption(self.key_file(user) if self.encrypted else None)
        return c

    def tail_logs(self, n: int = 30) -> str:
        out = []
        for f in sorted((self.home / "runtime" / "logs").glob("*.log")):
            lines = f.read_text(errors="replace").splitlines()[-n:]
            out.append(f"==> {f.name} <==\n" + "\n".join(lines))
        return "\n".join(out)

    def stop(self) -> None:
        from .cli.node import cli_node_stop
        from .cli.server import cli_server_stop

        CliRunner().invoke(cli_node_stop, ["--all"])
        CliRunner().invoke(cli_server_stop, ["--user", "--all"])

    def remove(self) -> None:
        """Stop everything and delete what ``create`` wrote (configurations, keys, entities, database, logs, runtime files)."""
        import shutil

        self.stop()
        for path in (self._cfg_dir("server") / f"{self.name}.yaml", self.fixtures_file if hasattr(self, "fixtures_file") else None,
                     self.description_file(), *[self._cfg_dir("node") / f"{self.name}-node-{i}.yaml" for i in range(self.n_nodes)]):
            if path is not None and Path(path).exists():
                Path(path).unlink()
        for sub in ("keys", "tasks"):
            shutil.rmtree(self.home / sub, ignore_errors=True)
        for kind in ("server", "node"):
            for scope in ("user", "system"):
                for base in ("data", "log"):
                    root = self.home / scope / base / kind
                    if root.is_dir():
                        for d in root.iterdir():
                            if d.name == self.name or d.name.startswith(f"{self.name}-node-"):
                                shutil.rmtree(d, ignore_errors=True)
