"""``vdev`` -- a whole demo network on this box with one command (the developer tool later vantage6 releases ship as
``vdev`` / ``v6 dev``; the 3.1 reference has no equivalent, its tests patch the docker SDK instead).

    vdev create-demo-network -n demo --nodes 8 --gpus 0,1,2,3,4,5,6,7     # configurations, keys, entities, database
    vdev start-demo-network  -n demo                                       # vserver start + N x vnode start
    vdev stop-demo-network   -n demo
    vdev remove-demo-network -n demo

Everything is driven through the real ``vserver`` / ``vnode`` commands (``vantage6_b200.dev.DemoNetwork``) and lives under
``--home`` (default ``$V6B200_HOME`` or ``./.v6b200``).  The researcher account is ``user-0`` / ``demo-password``.
"""
from __future__ import annotations

import click

from ..common import error, info
from ..dev import DemoNetwork


@click.group(name="vdev")
def cli_dev():
    """Create, start, stop and remove a demo network (server + nodes) on this machine."""


def _common(fn):
    fn = click.option("--home", default=None, help="root folder of the network (default: $V6B200_HOME or ./.v6b200)")(fn)
    return click.option("-n", "--name", default="demo", show_default=True, help="name of the network")(fn)


@cli_dev.command(name="create-demo-network")
@_common
@click.option("--nodes", "n_nodes", default=2, show_default=True, type=int, help="number of organizations / nodes")
@click.option("--gpus", default=None, help="comma-separated GPU indices, one per node (default: CPU nodes)")
@click.option("--database", "databases", multiple=True, help="database of node i (repeat per node; default: synthetic://node-i)")
@click.option("--encrypted/--not-encrypted", default=False, show_default=True, help="encrypted collaboration (RSA keys per organization)")
@click.option("--tls/--no-tls", default=False, show_default=True, help="https / wss with a self-signed certificate")
def create(name, home, n_nodes, gpus, databases, encrypted, tls):
    """Write the configurations, keys and entities of a network and import them into a fresh server database."""
    gpu_list = [int(g) for g in gpus.split(",")] if gpus else None
    if gpu_list is not None and len(gpu_list) != n_nodes:
        error(f"--gpus names {len(gpu_list)} devices for {n_nodes} nodes")
        raise SystemExit(1)
    if databases and len(databases) != n_nodes:
        error(f"--database given {len(databases)} times for {n_nodes} nodes")
        raise SystemExit(1)
    net = DemoNetwork(n_nodes, home=home, name=name, gpus=gpu_list, databases=list(databases) or None, encrypted=encrypted, tls=tls)
    net.create()
    info(f"Created demo network {name!r}: {n_nodes} node(s), server port {net.port}, description {net.description_file()}")
    info(f"Start it with `vdev start-demo-network -n {name}`; researcher login user-0 / {net.password}")


def _load(name, home) -> DemoNetwork:
    try:
        return DemoNetwork.load(name, home)
    except FileNotFoundError:
        error(f"No demo network {name!r} found: run `vdev create-demo-network -n {name}` first")
        raise SystemExit(1)


@cli_dev.command(name="start-demo-network")
@_common
def start(name, home):
    """Start the server and every node of a created network and wait until all nodes are online."""
    net = _load(name, home).up()
    info(f"Demo network {name!r} is up: {net.server_url}:{net.port}/api, collaboration id {net.collaboration_id}, "
         f"organizations {net.org_ids}")


@cli_dev.command(name="stop-demo-network")
@_common
def stop(name, home):
    """Stop the nodes and the server."""
    _load(name, home).stop()
    info(f"Demo network {name!r} stopped")


@cli_dev.command(name="remove-demo-network")
@_common
def remove(name, home):
    """Stop the network and delete its configurations, keys, database and logs."""
    _load(name, home).remove()
    info(f"Demo network {name!r} removed")


if __name__ == "__main__":
    cli_dev()
