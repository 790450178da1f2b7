"""What ``vnode`` and ``vserver`` have in common, written once.

A *managed instance* is a named configuration (in the user or system folders) that may be running
as a process of the local runtime (vantage6_b200/runtime).  Nodes and servers differ in their label,
the suffix of the runtime name and the noun used in console messages -- everything else (the
status table, stop / attach / version, the "new configuration" flow, log following) is the same
code, parameterised by an :class:`InstanceKind`.

The click commands in ``cli/node.py`` and ``cli/server.py`` pass *their own module* as ``deps``:
collaborators (``NodeContext``, ``q``, ``configuration_wizard``, ...) are looked up on it at call
time, so the unit tests can patch them where the reference's tests patch them
(reference tests/test_node_cli.py, tests/test_server_cli.py).

Console texts, option names and exit codes follow the reference CLI (SURVEY.md Appendix A;
reference vantage6/cli/node.py:71-119,428-502,734-763 and vantage6/cli/server.py:279-331,556-692).
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from threading import Thread
from typing import Callable, Dict, List, Optional, Tuple

import click

from ..common import error, info, warning
from ..common.colors import Fore, Style
from ..common.globals import APPNAME, STRING_ENCODING

ENVIRONMENTS = ("application", "prod", "acc", "test", "dev")
_COLUMNS = (("Name", 25), ("Environments", 32), ("Status", 16), ("System/User", 0))


def scope_of(system_folders: bool) -> str:
    return "system" if system_folders else "user"


def green(text) -> str:
    return f"{Fore.GREEN}{text}{Style.RESET_ALL}"


def red(text) -> str:
    return f"{Fore.RED}{text}{Style.RESET_ALL}"


@dataclass(frozen=True)
class InstanceKind:
    type_: str                    # "node" | "server": runtime label value and wizard type
    context_attr: str             # attribute of the CLI module holding the context class
    name_suffix: str              # appended to "{APPNAME}-{name}-{scope}"
    noun: str                     # "Node" | "Server" (capitalised, as printed)
    local_cli: str                # runtime-side CLI answering `version`

    @property
    def label(self) -> str:
        return f"{APPNAME}-type={self.type_}"

    @property
    def plural(self) -> str:
        return f"{self.type_}s"

    def runtime_name(self, name: str, system_folders: bool) -> str:
        return f"{APPNAME}-{name}-{scope_of(system_folders)}{self.name_suffix}"

    def running(self, rt) -> List[str]:
        return [c.name for c in rt.containers.list(filters={"label": self.label})]

    def context(self, deps):
        return getattr(deps, self.context_attr)


NODE = InstanceKind("node", "NodeContext", "", "Node", "vnode-local")
SERVER = InstanceKind("server", "ServerContext", "-server", "Server", "vserver-local")


# --------------------------------------------------------------------------------- click options
def folders_option(default: bool):
    """``--system`` / ``--user`` writing into ``system_folders``."""

    def wrap(fn):
        fn = click.option("--user", "system_folders", flag_value=False, default=default)(fn)
        return click.option("--system", "system_folders", flag_value=True)(fn)

    return wrap


def name_option(help_: Optional[str] = "configuration name"):
    return click.option("-n", "--name", default=None, help=help_)


def environment_option(default: str):
    return click.option("-e", "--environment", default=default, help="configuration environment to use")


def config_option():
    return click.option("-c", "--config", default=None, help="absolute path to configuration-file; overrides NAME")


# ----------------------------------------------------------------------------------------- guards
_NAME_CHARSET = "a-zA-Z0-9_.-"
_VALID_NAME = re.compile(f"[{_NAME_CHARSET}]+")


def check_config_name_allowed(name: str) -> None:
    """Configuration names end up in process, volume and file names: exit(1) unless ``name`` sticks to
    ``a-zA-Z0-9_.-`` (reference vantage6/cli/utils.py:6-11)."""
    if not name or _VALID_NAME.fullmatch(name) is None:
        error(f"Name '{name}' is not allowed. Please use only the following characters: {_NAME_CHARSET}")
        exit(1)


def check_if_docker_deamon_is_running(docker_client) -> None:
    """Exit(1) unless the local runtime answers a ping.  (The name is the reference's -- utils.py:14-19 --
    where the thing being pinged is the Docker daemon.)"""
    try:
        docker_client.ping()
    except Exception:  # noqa: BLE001
        error("Docker socket can not be found. Make sure Docker is running.")
        exit(1)


# ------------------------------------------------------------------------------------ log tailing
def print_log_worker(logs_stream) -> None:
    for chunk in logs_stream:
        print(chunk.decode(STRING_ENCODING), end="")


def follow_logs(deps, container, **attach_kw) -> None:
    """Stream an instance's log to the console until Ctrl-C."""
    stream = container.attach(stream=True, logs=True, **attach_kw)
    Thread(target=deps.print_log_worker, args=(stream,), daemon=True).start()
    try:
        while True:
            deps.time.sleep(1)
    except KeyboardInterrupt:
        info("Closing log file. Keyboard Interrupt.")
        exit(0)


# ------------------------------------------------------------------------------------ status table
def table_header() -> str:
    cells = []
    for title, width in _COLUMNS:
        cells.append(title.ljust(width) if width else title)
    return "\n" + "".join(cells)


def show_table(deps, kind: InstanceKind) -> None:
    """One row per configuration and folder scope, Online / Offline from the runtime's view."""
    rt = deps.docker.from_env()
    deps.check_docker_running()
    live = set(kind.running(rt))
    header = table_header()
    click.echo(header)
    click.echo("-" * len(header))
    broken = 0
    for system_folders, tag in ((True, " System "), (False, " User   ")):
        configs, failed = kind.context(deps).available_configurations(system_folders=system_folders)
        broken += len(failed)
        for cfg in configs:
            up = kind.runtime_name(cfg.name, system_folders) in live
            status = green("Online") if up else red("Offline")
            click.echo(f"{cfg.name:25}{str(cfg.available_environments):32}{status:25}{tag}")
    click.echo("-" * 85)
    if broken:
        warning(red(f"Failed imports: {broken}"))


# ---------------------------------------------------------------------- pick a running instance
def _resolve_running(deps, kind: InstanceKind, rt, name, system_folders, verb: str) -> Tuple[str, List[str]]:
    live = kind.running(rt)
    if name:
        return kind.runtime_name(name, system_folders), live
    picked = deps.q.select(f"Select the {kind.type_} you wish to {verb}:", choices=live).ask()
    return picked, live


def stop(deps, kind: InstanceKind, name, system_folders, everything: bool,
         halt: Callable[[object, str], None]) -> None:
    """``halt(rt, runtime_name)`` ends one instance (nodes: graceful stop, servers: kill + sidecar)."""
    rt = deps.docker.from_env()
    deps.check_docker_running()
    live = kind.running(rt)
    if not live:
        warning(f"No {kind.plural} are currently running.")
        return
    if everything:
        for runtime_name in live:
            halt(rt, runtime_name)
        return
    target, _ = _resolve_running(deps, kind, rt, name, system_folders, "stop")
    if target in live:
        halt(rt, target)
    else:
        shown = target if kind is NODE else name
        error(f"{red(shown)} is not running{'?' if kind is NODE else '!'}")


def attach(deps, kind: InstanceKind, name, system_folders, **attach_kw) -> None:
    rt = deps.docker.from_env()
    deps.check_docker_running()
    target, live = _resolve_running(deps, kind, rt, name, system_folders, "inspect")
    if target in live:
        follow_logs(deps, rt.containers.get(target), **attach_kw)
    else:
        error(f"{red(target)} was not running!?")


def version(deps, kind: InstanceKind, name, system_folders, cli_version: str) -> None:
    rt = deps.docker.from_env()
    deps.check_docker_running()
    live = kind.running(rt)
    if not name and not live:
        error(f"No {kind.plural} are running! You can only check the version for {kind.plural} that are running")
        exit(1)
    target, _ = _resolve_running(deps, kind, rt, name, system_folders, "inspect")
    if target not in live:
        error(f"{kind.noun} {target} is not running! Cannot provide version...")
        return
    reply = rt.containers.get(target).exec_run(cmd=f"{kind.local_cli} version", stdout=True)
    click.echo({kind.type_: reply.output.decode("utf-8"), "cli": cli_version})


# ------------------------------------------------------------------------------ new configuration
def ask_configuration_name(deps, name: Optional[str], always_normalise: bool) -> str:
    """Prompt when no name was given; spaces become dashes (and the user is told)."""
    prompted = not name
    if prompted:
        name = deps.q.text("Please enter a configuration-name:").ask()
    if prompted or always_normalise:
        dashed = name.replace(" ", "-")
        if dashed != name:
            info(f"Replaced spaces from configuration name: {dashed if always_normalise else name}")
            name = dashed
    return name


def require_write_access(deps, system_folders: bool, hint: Optional[str] = None) -> None:
    if deps.check_config_write_permissions(system_folders):
        return
    error("Your user does not have write access to all folders. Exiting")
    if hint:
        info(hint)
    exit(1)


def announce_new_configuration(kind: InstanceKind, cfg_file, flag: str) -> None:
    info(f"New configuration created: {green(cfg_file)}")
    info(f"You can start the {kind.type_} by running {Fore.GREEN}v{kind.type_} start {flag}{Style.RESET_ALL}")


# ----------------------------------------------------------------------------------- launch plans
@dataclass
class LaunchPlan:
    """Everything ``containers.run`` needs, collected step by step by the ``start`` commands."""

    image: Optional[str] = None
    command: str = ""
    name: Optional[str] = None
    labels: Dict[str, str] = field(default_factory=dict)
    environment: Dict[str, str] = field(default_factory=dict)
    binds: List[Tuple[str, str]] = field(default_factory=list)        # (target inside, source on the host)
    keep: bool = False

    def bind(self, target: str, source) -> None:
        self.binds.append((target, str(source)))

    def volume_specs(self) -> List[str]:
        return [f"{source}:{target}" for target, source in self.binds]


def choose_image(deps, ctx, requested: Optional[str], default_image: str, what: str, tail: str = "") -> str:
    """CLI option > configuration file > package default; then a best-effort refresh."""
    image = requested or ctx.config.get("image", default_image)
    info(f"Pulling latest {what} image '{image}'{tail}")
    return image


def report_pull(ok: bool, lead: str = " ") -> None:
    if ok:
        info(" ... success!")
    else:
        warning(f"{lead}... alas, no dice!")
