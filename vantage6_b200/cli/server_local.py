"""``vserver-local`` -- the server runtime's own entry point, i.e. what the reference starts
*inside* the server container (reference vantage6/cli/server.py:223-228 ``uwsgi ... --pyargv
/mnt/config.yaml``; ``vserver-local import`` :488; ``vserver-local shell`` :546-547;
``vserver-local version`` :687)."""
from __future__ import annotations

import code
import json
import os
import signal
import sys

import click
import yaml

from .. import __version__
from ..common import error, info
from .context import ServerContext
from .globals import DEFAULT_SERVER_ENVIRONMENT


def _ctx(config, environment, system_folders=True):
    return ServerContext.from_external_config_file(config, environment, system_folders)


def _make_app(ctx):
    from ..server.app import ServerApp

    cfg = dict(ctx.config)
    cfg["uri"] = ctx.get_database_uri()
    return ServerApp(cfg, data_dir=ctx.data_dir, name=ctx.name)


@click.group(name="vserver-local")
def cli_server_local():
    """Server runtime commands."""


@cli_server_local.command(name="start")
@click.option("-c", "--config", required=True)
@click.option("-e", "--environment", default=DEFAULT_SERVER_ENVIRONMENT)
@click.option("--ip", default=None)
@click.option("-p", "--port", default=None, type=int)
@click.option("--system", "system_folders", flag_value=True, default=True)
@click.option("--user", "system_folders", flag_value=False)
def start(config, environment, ip, port, system_folders):
    ctx = _ctx(config, environment, system_folders)
    app = _make_app(ctx)
    ports = json.loads(os.environ.get("V6_PORTS", "{}"))
    mapped = next(iter(ports.values()), None)
    if mapped:
        ip = ip or mapped[0]
        port = port or int(mapped[1])
    ip = ip or ctx.config.get("ip") or "127.0.0.1"
    port = port or int(ctx.config.get("port") or 5000)
    if ctx.config.get("rabbitmq_uri"):
        from ..server.mq_broker import attach_app

        attach_app(app, ctx.config["rabbitmq_uri"])
    signal.signal(signal.SIGTERM, lambda *_: sys.exit(0))
    scheme = "https" if (ctx.config.get("ssl") or {}).get("certfile") else "http"
    print(f"vantage6-b200 server '{ctx.name}' v{__version__} listening on {scheme}://{ip}:{port}{app.api_path}", flush=True)
    app.start(ip, port, block=True)


@cli_server_local.command(name="import")
@click.argument("file_", type=click.Path(exists=True))
@click.option("-c", "--config", required=True)
@click.option("-e", "--environment", default=DEFAULT_SERVER_ENVIRONMENT)
@click.option("--drop-all", is_flag=True, default=False)
def import_(file_, config, environment, drop_all):
    from ..server import fixtures

    ctx = _ctx(config, environment)
    app = _make_app(ctx)
    info("Reading yaml file.")
    with open(file_) as f:
        entities = yaml.safe_load(f.read())
    info("Adding entities to database.")
    try:
        counts = fixtures.load(app.db, entities, drop_all=drop_all, ensure_defaults=app.ensure_defaults)
    except PermissionError as e:
        error(str(e))
        sys.exit(1)
    info(f"Imported: {counts}")


@cli_server_local.command(name="shell")
@click.option("-c", "--config", required=True)
@click.option("-e", "--environment", default=DEFAULT_SERVER_ENVIRONMENT)
def shell(config, environment):
    """Interactive python shell with ``app`` and ``db`` bound (IPython is not required)."""
    ctx = _ctx(config, environment)
    app = _make_app(ctx)
    from ..server.admin_routes import issue_reset_token

    banner = f"vantage6-b200 server shell ({ctx.name}); objects: app, db, ctx; reset_token(username) mints a password reset token"
    code.interact(banner=banner, local={"app": app, "db": app.db, "ctx": ctx, "reset_token": lambda username: issue_reset_token(app, username)})


@cli_server_local.command(name="version")
def version():
    click.echo(__version__)


def main():
    argv = sys.argv[1:]
    # `uwsgi --http :5000 ... --pyargv /mnt/config.yaml` (the reference's launch string) maps to start
    if argv and argv[0] == "uwsgi":
        cfg = argv[argv.index("--pyargv") + 1] if "--pyargv" in argv else None
        port = None
        if "--http" in argv:
            port = int(argv[argv.index("--http") + 1].rsplit(":", 1)[-1])
        args = ["start", "-c", cfg]
        if port and not os.environ.get("V6_PORTS"):
            args += ["-p", str(port)]
        return cli_server_local.main(args=args, prog_name="vserver-local")
    return cli_server_local.main(prog_name="vserver-local")


if __name__ == "__main__":
    main()
