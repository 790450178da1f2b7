"""Interactive prompts with the slice of the ``questionary`` API the reference wizard uses
(reference vantage6/cli/configuration_wizard.py: ``q.prompt([...])``, ``q.text``, ``q.password``,
``q.select``, ``q.confirm``, ``q.Choice`` -- each returning an object with ``.ask()``).
questionary is not installed here; this is built on ``click.prompt`` so that it works on any
terminal and is trivially scriptable (``CliRunner(input=...)``) and mockable (tests patch the
module-level ``q``, exactly like the reference tests do: reference tests/test_wizard.py:18-39).
"""
from __future__ import annotations

from typing import Any, Dict, List, Sequence

import click


class _Question:
    def __init__(self, fn):
        self._fn = fn

    def ask(self):
        return self._fn()


class Choice:
    def __init__(self, title: str, value: Any = None):
        self.title = title
        self.value = title if value is None else value

    def __repr__(self):
        return f"Choice({self.title!r})"


def text(message: str, default: str = "", **_) -> _Question:
    return _Question(lambda: click.prompt(message, default=default, show_default=bool(default)))


def password(message: str, **_) -> _Question:
    return _Question(lambda: click.prompt(message, hide_input=True, default="", show_default=False))


def confirm(message: str, default: bool = True, **_) -> _Question:
    return _Question(lambda: click.confirm(message, default=default))


def select(message: str, choices: Sequence, **_) -> _Question:
    norm = [c if isinstance(c, Choice) else Choice(str(c), c) for c in choices]

    def run():
        click.echo(message)
        for i, c in enumerate(norm, 1):
            click.echo(f"  {i}) {c.title}")
        idx = click.prompt("Select", type=click.IntRange(1, len(norm)), default=1)
        return norm[idx - 1].value

    return _Question(run)


def prompt(questions: List[Dict[str, Any]], **_) -> Dict[str, Any]:
    """Ask a list of ``{"type","name","message"[,"default"][,"choices"]}`` questions."""
    answers: Dict[str, Any] = {}
    for qd in questions:
        kind = qd.get("type", "text")
        if kind == "text":
            answers[qd["name"]] = text(qd["message"], default=qd.get("default", "")).ask()
        elif kind == "password":
            answers[qd["name"]] = password(qd["message"]).ask()
        elif kind == "confirm":
            answers[qd["name"]] = confirm(qd["message"], default=qd.get("default", True)).ask()
        elif kind == "select":
            answers[qd["name"]] = select(qd["message"], qd["choices"]).ask()
        else:
            raise ValueError(f"unknown question type {kind!r}")
    return answers
