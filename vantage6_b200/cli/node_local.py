"""``vnode-local`` -- the node runtime's own entry point: what the reference starts *inside* the
node container with ``vnode-local start -c /mnt/config/{name}.yaml -n {name} -e {env}
--dockerized --system/--user`` (reference vantage6/cli/node.py:380-382) and queries with
``vnode-local version`` (reference node.py:759)."""
from __future__ import annotations

import signal
import sys

import click

from .. import __version__
from .context import NodeContext
from .globals import DEFAULT_NODE_ENVIRONMENT, DEFAULT_NODE_SYSTEM_FOLDERS


@click.group(name="vnode-local")
def cli_node_local():
    """Node runtime commands."""


@cli_node_local.command(name="start")
@click.option("-n", "--name", default=None)
@click.option("-c", "--config", default=None)
@click.option("-e", "--environment", default=DEFAULT_NODE_ENVIRONMENT)
@click.option("--system", "system_folders", flag_value=True)
@click.option("--user", "system_folders", flag_value=False, default=DEFAULT_NODE_SYSTEM_FOLDERS)
@click.option("--dockerized/--non-dockerized", default=False,
              help="accepted for command-line parity; nodes are processes pinned to a GPU here")
def start(name, config, environment, system_folders, dockerized):
    from ..node import Node

    if config:
        ctx = NodeContext.from_external_config_file(config, environment, system_folders)
        if name:
            ctx.name = name
    else:
        ctx = NodeContext(name, environment, system_folders)
    node = Node(ctx)

    def _term(*_):
        node.stop()
        sys.exit(0)

    signal.signal(signal.SIGTERM, _term)
    node.start(block=True)


@cli_node_local.command(name="version")
def version():
    click.echo(__version__)


def main():
    cli_node_local.main(prog_name="vnode-local")


if __name__ == "__main__":
    main()
