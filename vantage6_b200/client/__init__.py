"""Client library (the ``vantage6-client`` package the reference CLI depends on: reference
vantage6/cli/node.py:43,781-784 ``Client(host, port, api_path).authenticate(user, pw)``,
``.whoami.organization_name/.organization_id``, ``.request(path, method=, json=)``).

* ``ClientBase``     -- REST + JWT plumbing, token refresh, optional RSA encryption.
* ``UserClient``     -- researcher-facing API: ``client.task.create(...)``,
  ``client.wait_for_results(task_id)``, ``client.result.from_task(...)``, plus
  organization / collaboration / node / user / role / rule / util sub-clients.
  ``Client`` is an alias (vantage6 3.x name).
* ``ContainerClient``-- what a running algorithm's *master* uses (through the node's proxy) to
  create sub-tasks and collect their results.
"""
from __future__ import annotations

import json as _json
import logging
import time
from types import SimpleNamespace
from typing import Any, Dict, List, Optional, Union

from ..common import STRING_ENCODING, bytes_to_base64s  # noqa: F401
from ..common.encryption import CryptorBase, DummyCryptor, RSACryptor
from ..common.jsonhttp import JsonHttp
from ..common.serialization import deserialize, serialize

module_name = __name__.split(".")[-1]


class ServerError(Exception):
    def __init__(self, status: int, msg: str):
        super().__init__(f"[{status}] {msg}")
        self.status, self.msg = status, msg


def _wait_for_task(client: "ClientBase", task_id: int, sleep: float, timeout: float, include_results: bool = False) -> dict:
    """Shared by the researcher's and the algorithm's client: returns the task once it is complete (with its results when
    asked, so that no second request is needed).  The event cursor is taken BEFORE the first completeness check, so no
    wake-up can be missed; ``status_update`` events say whether they completed the task, so the task is only fetched
    again when one of them does (or when a poll times out, as a safety net)."""
    t0 = time.time()
    params = {"include": "results"} if include_results else None
    try:
        cursor = client.request("event", params={"timeout": 0}, timeout=15).get("last_id")
    except Exception:  # noqa: BLE001 -- no event channel: plain polling
        cursor = None
    while True:
        task = client.request(f"task/{task_id}", params=params)
        if task.get("complete"):
            return task
        while True:
            if time.time() - t0 > timeout:
                raise TimeoutError(f"task {task_id} did not complete in {timeout}s")
            if cursor is None:
                time.sleep(sleep)
                break
            try:
                reply = client.request("event", params={"since": cursor, "timeout": 5, "task_id": task_id}, timeout=15)
                cursor = reply.get("last_id", cursor)
                evs = [e for e in reply.get("events", []) if e.get("name") == "status_update" and e.get("data", {}).get("task_id") == task_id]
                if not reply.get("events") or any(e["data"].get("task_complete", True) for e in evs):
                    break                       # timed out (safety re-check) or an event that may have completed the task
            except Exception:  # noqa: BLE001
                cursor = None
                break


class ClientBase:
    def __init__(self, host: str, port: Optional[int] = 5000, path: str = "/api", ca_file: Optional[str] = None):
        self.log = logging.getLogger(module_name)
        self.__host, self.__port, self.__api_path = host.rstrip("/"), port, path
        self._access_token: Optional[str] = None
        self.__refresh_token: Optional[str] = None
        self.__refresh_url: Optional[str] = None
        self.cryptor: Optional[CryptorBase] = None
        self.whoami: Optional[SimpleNamespace] = None
        self._http = JsonHttp(ca_file)           # keep-alive connections; never routed through a proxy; https verified against ca_file

    # -- addresses ---------------------------------------------------------------------------
    @property
    def host(self):
        return self.__host

    @property
    def port(self):
        return self.__port

    @property
    def path(self):
        return self.__api_path

    @property
    def base_path(self) -> str:
        if self.__port:
            return f"{self.host}:{self.port}{self.__api_path}"
        return f"{self.host}{self.__api_path}"

    def generate_path_to(self, endpoint: str) -> str:
        if endpoint.startswith("/"):
            return self.base_path + endpoint
        return self.base_path + "/" + endpoint

    @property
    def headers(self) -> Dict[str, str]:
        return {"Authorization": "Bearer " + self._access_token} if self._access_token else {}

    @property
    def token(self):
        return self._access_token

    # -- requests ----------------------------------------------------------------------------
    def request(self, endpoint: str, json: dict = None, method: str = "get", params: dict = None,
                first_try: bool = True, timeout: float = 70.0):
        url = self.generate_path_to(endpoint)
        verb = method.upper() if method.lower() in ("get", "post", "put", "patch", "delete") else "GET"
        response = self._http.request(verb, url, json=json, headers=self.headers, params=params, timeout=timeout)
        if response.status_code > 210:
            try:
                msg = response.json().get("msg", response.text)
            except Exception:  # noqa: BLE001
                msg = response.text
            if response.status_code == 401 and first_try and self.__refresh_token and "expired" in str(msg).lower():
                self.refresh_token()
                return self.request(endpoint, json, method, params, first_try=False)
            raise ServerError(response.status_code, str(msg))
        return response.json()

    # -- authentication ----------------------------------------------------------------------
    def authenticate(self, credentials: dict, path: str = "token/user") -> None:
        data = self.request(path, json=credentials, method="post")
        self._access_token = data.get("access_token")
        self.__refresh_token = data.get("refresh_token")
        self.__refresh_url = data.get("refresh_url")
        self._auth_reply = data

    def refresh_token(self) -> None:
        assert self.__refresh_token, "Refresh token not found, did you authenticate?"
        url = f"{self.host}:{self.port}{self.__refresh_url}" if self.port else f"{self.host}{self.__refresh_url}"
        r = self._http.request("POST", url, headers={"Authorization": "Bearer " + self.__refresh_token}, timeout=30)
        if r.status_code != 200:
            raise ServerError(r.status_code, "Authentication Error!")
        self._access_token = r.json()["access_token"]

    # -- encryption --------------------------------------------------------------------------
    def setup_encryption(self, private_key_file: Optional[str]) -> None:
        """``None`` disables encryption (un-encrypted collaboration)."""
        if private_key_file is None:
            self.cryptor = DummyCryptor()
        else:
            self.cryptor = RSACryptor(private_key_file)

    def _encrypt_input(self, blob: bytes, organization_id: int) -> str:
        assert self.cryptor is not None, "Encryption has not yet been setup! (call setup_encryption)"
        if isinstance(self.cryptor, RSACryptor):
            org = self.request(f"organization/{organization_id}")
            pub = org.get("public_key")
            if not pub:
                raise ValueError(f"organization {organization_id} has no public key; cannot encrypt its input")
            return self.cryptor.encrypt_bytes_to_str(blob, pub)
        return self.cryptor.encrypt_bytes_to_str(blob, "")

    def _decrypt_result(self, value: Optional[str]) -> Any:
        if value is None or value == "":
            return None
        assert self.cryptor is not None, "Encryption has not yet been setup! (call setup_encryption)"
        try:
            return deserialize(self.cryptor.decrypt_str_to_bytes(value))
        except Exception:  # noqa: BLE001 -- not one of ours (e.g. plain JSON from a fixture)
            return value


class UserClient(ClientBase):
    """Researcher client."""

    def __init__(self, host: str = "http://localhost", port: Optional[int] = 5000, path: str = "/api",
                 verbose: bool = False, ca_file: Optional[str] = None):
        super().__init__(host, port, path, ca_file=ca_file)
        self.util = self.Util(self)
        self.collaboration = self.Collaboration(self)
        self.organization = self.Organization(self)
        self.user = self.User(self)
        self.result = self.Result(self)
        self.task = self.Task(self)
        self.role = self.Role(self)
        self.node = self.Node(self)
        self.rule = self.Rule(self)

    def authenticate(self, username: str, password: str) -> None:  # type: ignore[override]
        super().authenticate({"username": username, "password": password}, path="token/user")
        user = self.request(self._auth_reply["user_url"][len(self.path):])
        org = self.request(f"organization/{user['organization']['id']}")
        self.whoami = SimpleNamespace(type_="user", id_=user["id"], name=user["username"],
                                      organization_id=org["id"], organization_name=org["name"])
        self.log.info("Successfully authenticated as %s (organization %s)", user["username"], org["name"])

    def wait_for_results(self, task_id: int, sleep: float = 0.05, timeout: float = 600.0) -> List[Any]:
        """Block until the task is complete, then return the decrypted results.  Woken by the server's
        ``status_update`` events (long poll); plain polling every ``sleep`` seconds is the fallback."""
        task = _wait_for_task(self, task_id, sleep, timeout, include_results=True)
        rows = task["results"]
        for r in rows:
            r["result"] = self._decrypt_result(r.get("result"))
        return rows

    def run(self, image: str, input: dict, collaboration: Optional[int] = None, organizations: Optional[List[int]] = None,  # noqa: A002
            name: str = "task", database: str = "default", timeout: float = 600.0, raise_on_failure: bool = True) -> List[Any]:
        """Create a task, wait for it and return the decoded results -- the three calls every script makes.  Defaults: the
        only collaboration you are in, your own organization (where a ``master`` usually runs).  A result that came back
        empty (failed algorithm, refused image) raises with the node's log unless ``raise_on_failure=False``."""
        if collaboration is None:
            mine = self.collaboration.list()
            if len(mine) != 1:
                raise ValueError(f"name the collaboration: you are in {len(mine)} of them")
            collaboration = mine[0]["id"]
        orgs = list(organizations) if organizations else [self.whoami.organization_id]
        task = self.task.create(collaboration=collaboration, organizations=orgs, name=name, image=image, input=input, database=database)
        rows = self.wait_for_results(task["id"], timeout=timeout)
        failed = [r for r in rows if r.get("result") is None]
        if failed and raise_on_failure:
            r = failed[0]
            raise RuntimeError(f"task {task['id']}: {len(failed)} of {len(rows)} results are empty (status {r.get('status')!r}, organization "
                               f"{r['organization']['id']}); log of the first:\n{(r.get('log') or '')[-2000:]}")
        return [r["result"] for r in rows]

    # ---------------------------------------------------------------- sub clients
    class SubClient:
        def __init__(self, parent: "UserClient"):
            self.parent = parent

    class Util(SubClient):
        def get_server_version(self) -> dict:
            return self.parent.request("version")

        def get_server_health(self) -> dict:
            return self.parent.request("health")

        def change_my_password(self, current_password: str, new_password: str) -> dict:
            return self.parent.request("password/change", method="patch",
                                       json={"current_password": current_password, "new_password": new_password})

        def reset_my_password(self, email: Optional[str] = None, username: Optional[str] = None) -> dict:
            """Ask for a reset token (mailed, or handed out by the server's operator); no authentication needed."""
            assert email or username, "You need to provide username or email!"
            return self.parent.request("recover/lost", method="post", json={"username": username, "email": email})

        def set_my_password(self, token: str, password: str) -> dict:
            return self.parent.request("recover/reset", method="post", json={"reset_token": token, "password": password})

    class Collaboration(SubClient):
        def list(self) -> List[dict]:
            return self.parent.request("collaboration")

        def get(self, id_: int) -> dict:
            return self.parent.request(f"collaboration/{id_}")

        def create(self, name: str, organizations: List[int], encrypted: bool = False) -> dict:
            return self.parent.request("collaboration", method="post",
                                       json={"name": name, "organization_ids": organizations, "encrypted": encrypted})

        def update(self, id_: int, **fields) -> dict:
            return self.parent.request(f"collaboration/{id_}", method="patch", json=fields)

        def delete(self, id_: int) -> dict:
            return self.parent.request(f"collaboration/{id_}", method="delete")

        def organizations(self, id_: int) -> List[dict]:
            return self.parent.request(f"collaboration/{id_}/organization")

        def nodes(self, id_: int) -> List[dict]:
            return self.parent.request(f"collaboration/{id_}/node")

        def tasks(self, id_: int) -> List[dict]:
            return self.parent.request(f"collaboration/{id_}/task")

        def add_organization(self, id_: int, organization: int) -> List[dict]:
            return self.parent.request(f"collaboration/{id_}/organization", method="post", json={"id": organization})

        def remove_organization(self, id_: int, organization: int) -> List[dict]:
            return self.parent.request(f"collaboration/{id_}/organization", method="delete", json={"id": organization})

        def add_node(self, id_: int, node: int) -> List[dict]:
            return self.parent.request(f"collaboration/{id_}/node", method="post", json={"id": node})

        def remove_node(self, id_: int, node: int) -> List[dict]:
            return self.parent.request(f"collaboration/{id_}/node", method="delete", json={"id": node})

    class Organization(SubClient):
        def list(self) -> List[dict]:
            return self.parent.request("organization")

        def get(self, id_: Optional[int] = None) -> dict:
            id_ = id_ if id_ is not None else self.parent.whoami.organization_id
            return self.parent.request(f"organization/{id_}")

        def create(self, name: str, **fields) -> dict:
            return self.parent.request("organization", method="post", json={"name": name, **fields})

        def update(self, id_: Optional[int] = None, **fields) -> dict:
            id_ = id_ if id_ is not None else self.parent.whoami.organization_id
            return self.parent.request(f"organization/{id_}", method="patch", json=fields)

        def collaborations(self, id_: Optional[int] = None) -> List[dict]:
            id_ = id_ if id_ is not None else self.parent.whoami.organization_id
            return self.parent.request(f"organization/{id_}/collaboration")

        def nodes(self, id_: Optional[int] = None) -> List[dict]:
            id_ = id_ if id_ is not None else self.parent.whoami.organization_id
            return self.parent.request(f"organization/{id_}/node")

    class User(SubClient):
        def list(self) -> List[dict]:
            return self.parent.request("user")

        def get(self, id_: Optional[int] = None) -> dict:
            id_ = id_ if id_ is not None else self.parent.whoami.id_
            return self.parent.request(f"user/{id_}")

        def create(self, username: str, password: str, organization: Optional[int] = None, roles: List[int] = (),
                   rules: List[int] = (), **fields) -> dict:
            body = {"username": username, "password": password, "organization_id": organization, "roles": list(roles),
                    "rules": list(rules), **fields}
            return self.parent.request("user", method="post", json=body)

        def update(self, id_: Optional[int] = None, **fields) -> dict:
            id_ = id_ if id_ is not None else self.parent.whoami.id_
            return self.parent.request(f"user/{id_}", method="patch", json=fields)

        def delete(self, id_: int) -> dict:
            return self.parent.request(f"user/{id_}", method="delete")

    class Role(SubClient):
        def list(self) -> List[dict]:
            return self.parent.request("role")

        def get(self, id_: int) -> dict:
            return self.parent.request(f"role/{id_}")

        def create(self, name: str, description: str = "", rules: List[int] = (), organization: Optional[int] = None) -> dict:
            """A role in ``organization`` (default: your own) holding ``rules`` -- only rules you hold yourself."""
            body = {"name": name, "description": description, "rules": list(rules)}
            if organization is not None:
                body["organization_id"] = organization
            return self.parent.request("role", method="post", json=body)

        def update(self, id_: int, **fields) -> dict:
            """``name``, ``description`` and / or ``rules`` (the full new list of rule ids)."""
            return self.parent.request(f"role/{id_}", method="patch", json=fields)

        def delete(self, id_: int, delete_dependents: bool = False) -> dict:
            return self.parent.request(f"role/{id_}", method="delete",
                                       params={"delete_dependents": "true"} if delete_dependents else None)

        def rules(self, id_: int) -> List[dict]:
            return self.parent.request(f"role/{id_}/rule")

        def add_rule(self, id_: int, rule: int) -> dict:
            return self.parent.request(f"role/{id_}/rule/{rule}", method="post")

        def remove_rule(self, id_: int, rule: int) -> dict:
            return self.parent.request(f"role/{id_}/rule/{rule}", method="delete")

    class Rule(SubClient):
        def list(self) -> List[dict]:
            return self.parent.request("rule")

        def get(self, id_: int) -> dict:
            return self.parent.request(f"rule/{id_}")

    class Node(SubClient):
        def list(self) -> List[dict]:
            return self.parent.request("node")

        def get(self, id_: int) -> dict:
            return self.parent.request(f"node/{id_}")

        def create(self, collaboration: int, organization: Optional[int] = None, name: Optional[str] = None) -> dict:
            """Returns the node *including its api_key* (shown only once)."""
            return self.parent.request("node", method="post",
                                       json={"collaboration_id": collaboration, "organization_id": organization, "name": name})

        def update(self, id_: int, **fields) -> dict:
            return self.parent.request(f"node/{id_}", method="patch", json=fields)

        def delete(self, id_: int) -> dict:
            return self.parent.request(f"node/{id_}", method="delete")

        def tasks(self, id_: int, open_only: bool = False) -> List[dict]:
            return self.parent.request(f"node/{id_}/task", params={"state": "open"} if open_only else None)

    class Task(SubClient):
        def list(self, **filters) -> List[dict]:
            return self.parent.request("task", params=filters or None)

        def get(self, id_: int, include_results: bool = False) -> dict:
            return self.parent.request(f"task/{id_}", params={"include": "results"} if include_results else None)

        def create(self, collaboration: int, organizations: List[int], name: str, image: str, description: str = "",
                   input: dict = None, data_format: str = "json", database: str = "default") -> dict:  # noqa: A002
            """Create a task for ``organizations``; the serialized input is encrypted per
            receiving organization when the collaboration is encrypted."""
            assert self.parent.cryptor, "Encryption has not yet been setup!"
            blob = serialize(input or {}, data_format)
            orgs = [{"id": oid, "input": self.parent._encrypt_input(blob, oid)} for oid in organizations]
            return self.parent.request("task", method="post", json={
                "name": name, "image": image, "collaboration_id": collaboration, "description": description,
                "organizations": orgs, "database": database})

        def delete(self, id_: int) -> dict:
            return self.parent.request(f"task/{id_}", method="delete")

    class Result(SubClient):
        def get(self, id_: int) -> dict:
            r = self.parent.request(f"result/{id_}")
            r["result"] = self.parent._decrypt_result(r.get("result"))
            return r

        def list(self, **filters) -> List[dict]:
            return self.parent.request("result", params=filters or None)

        def from_task(self, task_id: int) -> List[dict]:
            rows = self.parent.request(f"task/{task_id}/result")
            for r in rows:
                r["result"] = self.parent._decrypt_result(r.get("result"))
            return rows


Client = UserClient


class ContainerClient(ClientBase):
    """Client used by an algorithm's *master* function.  It talks to the node's proxy server
    (host/port from the environment the node prepared), which adds encryption and forwards to
    the central server with the container token."""

    def __init__(self, token: str, host: str, port: Optional[int], path: str = ""):
        super().__init__(host, port, path)
        self._access_token = token
        import jwt

        claims = _json.loads(jwt.decode(token, options={"verify_signature": False})["sub"])
        self.image = claims.get("image")
        self.host_node_id = claims.get("node_id")
        self.collaboration_id = claims.get("collaboration_id")
        self.organization_id = claims.get("organization_id")
        self.task_id = claims.get("task_id")
        self.cryptor = DummyCryptor()     # the proxy handles the real encryption
        self.log.info("Container client for task %s in collaboration %s", self.task_id, self.collaboration_id)

    def authenticate(self, *a, **k):  # type: ignore[override]
        self.log.warning("Containers do not need to authenticate!")

    def refresh_token(self):  # type: ignore[override]
        self.log.warning("Containers cannot refresh their token!")

    def get_results(self, task_id: int) -> List[Any]:
        """Decoded outputs of all (finished) results of ``task_id``."""
        rows = self.request(f"task/{task_id}/result")
        out = []
        for r in rows:
            if r.get("result"):
                out.append(deserialize(self.cryptor.str_to_bytes(r["result"])))
        return out

    def get_task(self, task_id: int) -> dict:
        return self.request(f"task/{task_id}")

    def create_new_task(self, input_: dict, organization_ids: List[int] = (), name: str = "subtask",
                        description: Optional[str] = None, data_format: str = "json") -> dict:
        self.log.debug("Creating new subtask for organizations %s", list(organization_ids))
        blob = self.cryptor.bytes_to_str(serialize(input_, data_format))
        orgs = [{"id": oid, "input": blob} for oid in organization_ids]
        return self.request("task", method="post", json={
            "name": name, "image": self.image, "collaboration_id": self.collaboration_id,
            "description": description or f"task from container on node_id={self.host_node_id}", "organizations": orgs})

    def wait_for_task(self, task_id: int, sleep: float = 0.05, timeout: float = 3600.0) -> dict:
        """Block until ``task_id`` is complete; woken by the server's ``status_update`` events (long poll through the
        node's proxy) instead of sleeping between polls, with plain polling as the fallback."""
        return _wait_for_task(self, task_id, sleep, timeout)

    def wait_for_results(self, task_id: int, sleep: float = 0.05, timeout: float = 3600.0) -> List[Any]:
        """Wait for ``task_id`` and return the decoded outputs of its (non-empty) results -- the completeness check that
        ends the wait already carries them, so there is no second request."""
        task = _wait_for_task(self, task_id, sleep, timeout, include_results=True)
        return [deserialize(self.cryptor.str_to_bytes(r["result"])) for r in task.get("results", [])
                if isinstance(r, dict) and r.get("result")]

    def get_organizations_in_my_collaboration(self) -> List[dict]:
        return self.request(f"collaboration/{self.collaboration_id}/organization")

    def get_algorithm_addresses(self, task_id: int) -> List[dict]:
        """vantage6's VPN address book; on one NVSwitch box node-to-node traffic is symmetric
        memory, so the 'address' of a peer algorithm is its rank in the rendezvous.  Ports that the nodes registered
        for the algorithms of ``task_id`` (``/port``) are listed with the node they belong to."""
        nodes = self.request(f"collaboration/{self.collaboration_id}/node")
        try:
            ports = self.request("port", params={"task_id": task_id})
        except ServerError:
            ports = []
        out = []
        for i, n in enumerate(sorted(nodes, key=lambda n: n["id"])):
            mine = [p for p in ports if p.get("node_id") == n["id"]]
            out.append({"rank": i, "node_id": n["id"], "organization_id": n["organization"]["id"], "gpu": n.get("gpu"),
                        "ip": n.get("ip"), "ports": [{"port": p["port"], "label": p["label"]} for p in mine]})
        return out


__all__ = ["ClientBase", "UserClient", "Client", "ContainerClient", "ServerError", "RSACryptor", "DummyCryptor"]
