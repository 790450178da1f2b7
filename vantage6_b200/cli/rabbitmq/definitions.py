"""Boot-time definitions of the message-queue sidecar.

``RABBITMQ_DEFINITIONS`` is the document the queue manager fills in (``{{username}}``, ``{{password}}`` -- a
salted-SHA256 hash --, ``{{vhost_name}}``) and writes as ``definitions.json``: one administrator, one vhost, full
permissions on it, nothing pre-declared.  Same document shape as the reference
(vantage6/cli/rabbitmq/definitions.py:1-30), so a ``definitions.json`` written by either is readable by the in-box
broker (server/mq_broker.py).
"""

_EVERYTHING = ".*"


def _administrator():
    return {"name": "{{username}}", "password_hash": "{{password}}", "tags": "administrator",
            "hashing_algorithm": "rabbit_password_hashing_sha256"}


def _full_access():
    grant = {right: _EVERYTHING for right in ("configure", "write", "read")}
    return {"user": "{{username}}", "vhost": "{{vhost_name}}", **grant}


def build_definitions() -> dict:
    doc = {"rabbit_version": "3.6.6", "users": [_administrator()], "vhosts": [{"name": "{{vhost_name}}"}],
           "permissions": [_full_access()]}
    doc.update({section: [] for section in ("parameters", "policies", "queues", "exchanges", "bindings")})
    return doc


RABBITMQ_DEFINITIONS = build_definitions()
