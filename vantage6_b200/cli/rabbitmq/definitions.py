"""Start-up definitions of the message-queue sidecar (reference
vantage6/cli/rabbitmq/definitions.py:1-30): one administrator user with a salted-SHA256
password hash, one vhost, full permissions.  The in-box broker (server/mq_broker.py) reads the
same document, so a reference ``definitions.json`` stays meaningful."""

RABBITMQ_DEFINITIONS = {
    "rabbit_version": "3.6.6",
    "users": [{"name": "{{username}}", "password_hash": "{{password}}",
               "hashing_algorithm": "rabbit_password_hashing_sha256", "tags": "administrator"}],
    "vhosts": [{"name": "{{vhost_name}}"}],
    "permissions": [{"user": "{{username}}", "vhost": "{{vhost_name}}", "configure": ".*", "write": ".*", "read": ".*"}],
    "parameters": [], "policies": [], "queues": [], "exchanges": [], "bindings": [],
}
