"""Multi-environment YAML configuration files.

One file ``<config_dir>/<name>.yaml`` holds an ``application`` config and/or several
``environments`` (``prod``, ``acc``, ``test``, ``dev``); the wizard merges a new environment into
an existing file (reference vantage6/cli/configuration_wizard.py:234-242; usage of
``ConfigurationManager(conf_class, name)``, ``from_file``, ``put``, ``save`` at
reference vantage6/cli/configuration_manager.py:75-92 and vantage6/cli/node.py:596-597).
"""
from __future__ import annotations

import collections
from pathlib import Path
from typing import Dict, List, Type

import yaml

from .schema import Schema, SchemaError

ENVIRONMENTS = ("prod", "acc", "test", "dev")


class Configuration(collections.UserDict):
    """A dict validated against the class-level ``VALIDATORS`` spec."""

    VALIDATORS: dict = {}

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)

    @property
    def is_valid(self) -> bool:
        return Schema(self.VALIDATORS).is_valid(self.data)

    def validation_error(self) -> str | None:
        try:
            Schema(self.VALIDATORS).validate(self.data)
            return None
        except SchemaError as e:
            return str(e)

    def normalized(self) -> dict:
        """Data with ``Use(...)`` coercions applied (e.g. ``port: "5000"`` -> 5000)."""
        return Schema(self.VALIDATORS).validate(self.data)


class ConfigurationManager:
    """Class to maintain valid configuration settings of one instance (all environments)."""

    def __init__(self, conf_class: Type[Configuration] = Configuration, name: str | None = None):
        self.application: Configuration | dict = {}
        self.prod: Configuration | dict = {}
        self.acc: Configuration | dict = {}
        self.test: Configuration | dict = {}
        self.dev: Configuration | dict = {}
        self.name = name
        self.conf_class = conf_class

    # -- mutation ---------------------------------------------------------------------------
    def put(self, env: str, config: dict) -> None:
        assert env in ("application",) + ENVIRONMENTS, f"unknown environment {env!r}"
        cfg = self.conf_class(config)
        err = cfg.validation_error()
        if err is not None:
            raise SchemaError(f"invalid configuration for environment {env!r}: {err}")
        setattr(self, env, cfg)

    def get(self, env: str):
        assert env in ("application",) + ENVIRONMENTS, f"unknown environment {env!r}"
        return getattr(self, env)

    # -- inspection -------------------------------------------------------------------------
    @property
    def is_empty(self) -> bool:
        return not (self.application or self.prod or self.acc or self.test or self.dev)

    @property
    def environments(self) -> Dict[str, dict]:
        return {"prod": self.prod, "acc": self.acc, "test": self.test, "dev": self.dev}

    @property
    def has_application(self) -> bool:
        return bool(self.application)

    @property
    def has_environments(self) -> bool:
        return any(bool(v) for v in self.environments.values())

    @property
    def available_environments(self) -> List[str]:
        out = ["application"] if self.has_application else []
        out += [k for k, v in self.environments.items() if v]
        return out

    # -- persistence ------------------------------------------------------------------------
    def load(self, path) -> None:
        with open(str(path), "r") as f:
            config = yaml.safe_load(f) or {}
        if config.get("application"):
            self.put("application", config["application"])
        for env, cfg in (config.get("environments") or {}).items():
            if cfg:
                self.put(env, cfg)

    @classmethod
    def from_file(cls, path, conf_class: Type[Configuration] = Configuration) -> "ConfigurationManager":
        name = Path(path).stem
        assert name, f"could not derive a configuration name from {path!r}"
        conf = cls(name=name, conf_class=conf_class)
        conf.load(path)
        return conf

    def save(self, path) -> None:
        def plain(c):
            return dict(c.data) if isinstance(c, Configuration) else dict(c)

        config = {"application": plain(self.application),
                  "environments": {k: plain(v) for k, v in self.environments.items()}}
        Path(path).parent.mkdir(parents=True, exist_ok=True)
        with open(str(path), "w") as f:
            yaml.safe_dump(config, f, default_flow_style=False, sort_keys=False)
