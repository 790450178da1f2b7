"""Console helpers and small utilities -- the ``vantage6.common`` surface the reference CLI
imports (reference vantage6/cli/node.py:27-30, vantage6/cli/server.py:14-15, utils.py:3).

Output format is asserted by the reference tests: ``"[info]  - <msg>"`` / ``"[error] - <msg>"``
(reference tests/test_node_cli.py:105,122,250)."""
from __future__ import annotations

import base64
import os
from pathlib import Path

import click

from .colors import Fore, Style
from .globals import APPNAME, STRING_ENCODING  # noqa: F401


def logger_name(special_name: str) -> str:
    log_name = special_name.split(".")[-1]
    if len(log_name) > 14:
        log_name = log_name[:11] + ".."
    return log_name


class Singleton(type):
    _instances: dict = {}

    def __call__(cls, *args, **kwargs):
        if cls not in cls._instances:
            cls._instances[cls] = super().__call__(*args, **kwargs)
        return cls._instances[cls]


def bytes_to_base64s(bytes_: bytes) -> str:
    return base64.b64encode(bytes_).decode(STRING_ENCODING)


def base64s_to_bytes(bytes_string: str) -> bytes:
    return base64.b64decode(bytes_string.encode(STRING_ENCODING))


def echo(msg: str, level: str = "info") -> None:
    fmt = {
        "error": f"[{Fore.RED}error{Style.RESET_ALL}]",
        "warn": f"[{Fore.YELLOW}warn{Style.RESET_ALL}] ",
        "info": f"[{Fore.GREEN}info{Style.RESET_ALL}] ",
        "debug": f"[{Fore.CYAN}debug{Style.RESET_ALL}]",
    }
    click.echo(f"{Style.RESET_ALL}{fmt[level]} - {msg}")


def info(msg: str) -> None:
    echo(msg, "info")


def warning(msg: str) -> None:
    echo(msg, "warn")


def error(msg: str) -> None:
    echo(msg, "error")


def debug(msg: str) -> None:
    echo(msg, "debug")


def check_config_write_permissions(system_folders: bool = False) -> bool:
    """True if the current user can write the configuration / data / log folders."""
    from .context import AppContext

    dirs = AppContext.type_folders("node", system_folders)
    for d in dirs.values():
        p = Path(d)
        while not p.exists() and p != p.parent:
            p = p.parent
        if not os.access(p, os.W_OK):
            return False
    return True
