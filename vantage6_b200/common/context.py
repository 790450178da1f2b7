"""AppContext: where an instance's configuration, data and log files live, plus logging setup.

Re-implements the ``vantage6.common.context.AppContext`` contract that the reference CLI
subclasses (reference vantage6/cli/context.py:16-144): constructor
``(instance_type, instance_name, environment, system_folders[, config_file])``, attributes
``log, config, config_file, config_file_name, config_dir, config_manager, data_dir, log_dir,
log_file, name, scope, environment``, ``get_data_file()``, and the classmethods
``from_external_config_file``, ``config_exists``, ``available_configurations``,
``instance_folders``, ``type_data_folder``.

Folder policy: ``--system`` -> /etc/vantage6, /var/lib/vantage6, /var/log/vantage6;
``--user`` -> XDG dirs under the home directory.  ``$V6B200_HOME`` relocates everything under
one root (tests, sandboxes, the GPU box).
"""
from __future__ import annotations

import logging
import logging.handlers
import os
import sys
from pathlib import Path
from typing import Dict, List, Tuple

from . import logger_name
from .configuration_manager import ConfigurationManager
from .globals import APPNAME, HOME_ENV


class AppContext:
    INST_CONFIG_MANAGER = ConfigurationManager
    LOGGING_ENABLED = True

    def __init__(self, instance_type: str, instance_name: str, environment: str = "application",
                 system_folders: bool = False, config_file: str | Path | None = None):
        self.scope = "system" if system_folders else "user"
        self.name = instance_name
        self.instance_type = instance_type
        self.environment = environment
        self.config_manager = None
        self.config_file = None
        self.initialize(instance_type, instance_name, environment, system_folders, config_file)
        if self.LOGGING_ENABLED:
            self.set_folders(instance_type, self.name, system_folders)
            self.setup_logging()
        else:
            self.set_folders(instance_type, self.name, system_folders)
            self.log = logging.getLogger(logger_name(__name__))

    # ------------------------------------------------------------------ construction
    def initialize(self, instance_type, instance_name, environment, system_folders, config_file=None):
        self.name = instance_name
        if config_file is None:
            config_file = self.find_config_file(instance_type, instance_name, system_folders)
        self.config_file = Path(config_file)
        self.config_manager = self.INST_CONFIG_MANAGER.from_file(self.config_file)
        cfg = self.config_manager.get(environment)
        if not cfg:
            raise ValueError(f"environment {environment!r} is not present in {self.config_file}")
        self.config = cfg

    @classmethod
    def from_external_config_file(cls, path, instance_type, environment="application", system_folders=False):
        instance_name = Path(path).stem
        self_ = cls.__new__(cls)
        self_.scope = "system" if system_folders else "user"
        self_.instance_type = instance_type
        self_.environment = environment
        self_.name = instance_name
        self_.initialize(instance_type, instance_name, environment, system_folders, path)
        self_.set_folders(instance_type, instance_name, system_folders)
        if cls.LOGGING_ENABLED:
            self_.setup_logging()
        else:
            self_.log = logging.getLogger(logger_name(__name__))
        return self_

    # ------------------------------------------------------------------ folders
    @staticmethod
    def _roots(system_folders: bool) -> Dict[str, Path]:
        home = os.environ.get(HOME_ENV)
        if home:
            base = Path(home) / ("system" if system_folders else "user")
            return {"config": base / "config", "data": base / "data", "log": base / "log"}
        if system_folders:
            return {"config": Path("/etc") / APPNAME, "data": Path("/var/lib") / APPNAME,
                    "log": Path("/var/log") / APPNAME}
        h = Path.home()
        return {"config": Path(os.environ.get("XDG_CONFIG_HOME", h / ".config")) / APPNAME,
                "data": Path(os.environ.get("XDG_DATA_HOME", h / ".local" / "share")) / APPNAME,
                "log": Path(os.environ.get("XDG_CACHE_HOME", h / ".cache")) / APPNAME / "log"}

    @classmethod
    def type_folders(cls, instance_type: str, system_folders: bool) -> Dict[str, Path]:
        r = cls._roots(system_folders)
        return {k: v / instance_type for k, v in r.items()}

    @classmethod
    def type_data_folder(cls, instance_type: str, system_folders: bool = False) -> Path:
        return cls.type_folders(instance_type, system_folders)["data"]

    @classmethod
    def instance_folders(cls, instance_type: str, instance_name: str, system_folders: bool) -> Dict[str, Path]:
        d = cls.type_folders(instance_type, system_folders)
        return {"log": d["log"] / instance_name, "data": d["data"] / instance_name, "config": d["config"]}

    def set_folders(self, instance_type, instance_name, system_folders) -> None:
        dirs = self.instance_folders(instance_type, instance_name, system_folders)
        self.log_dir = dirs["log"]
        self.data_dir = dirs["data"]
        self.config_dir = dirs["config"]

    @classmethod
    def find_config_file(cls, instance_type, instance_name, system_folders, config_file=None) -> str:
        if config_file is None:
            config_file = f"{instance_name}.yaml"
        config_dir = cls.instance_folders(instance_type, instance_name, system_folders)["config"]
        full = Path(config_dir) / config_file
        if full.exists():
            return str(full)
        raise FileNotFoundError(f"could not find configuration file {config_file!r} in {config_dir}")

    @classmethod
    def config_exists(cls, instance_type, instance_name, environment="application", system_folders=False) -> bool:
        try:
            f = cls.find_config_file(instance_type, instance_name, system_folders)
        except Exception:  # noqa: BLE001
            return False
        try:
            cm = cls.INST_CONFIG_MANAGER.from_file(f)
        except Exception:  # noqa: BLE001
            return False
        return bool(cm.get(environment))

    @classmethod
    def available_configurations(cls, instance_type, system_folders) -> Tuple[List, List]:
        """Return ``(configs, failed)``: config managers of every ``*.yaml`` in the type's config
        folder and the files that failed to load (reference vantage6/cli/node.py:93-119)."""
        folder = cls.type_folders(instance_type, system_folders)["config"]
        configs, failed = [], []
        if not Path(folder).exists():
            return configs, failed
        for file_ in sorted(Path(folder).glob("*.yaml")):
            try:
                conf = cls.INST_CONFIG_MANAGER.from_file(file_)
                if conf.is_empty:
                    failed.append(file_)
                else:
                    configs.append(conf)
            except Exception:  # noqa: BLE001
                failed.append(file_)
        return configs, failed

    # ------------------------------------------------------------------ properties
    @property
    def config_file_name(self) -> str:
        return Path(self.config_file).stem

    @property
    def log_file(self) -> Path:
        assert self.config_manager, "log file unknown without a configuration"
        file_ = f"{self.config_manager.name}-{self.scope}.log"
        return self.log_dir / file_

    def get_data_file(self, filename: str) -> str:
        """Absolute paths are returned as-is; relative ones resolve against ``data_dir``."""
        if Path(filename).is_absolute():
            return str(filename)
        return str(self.data_dir / filename)

    # ------------------------------------------------------------------ logging
    def setup_logging(self) -> None:
        """Rotating file + optional console logging from the config's ``logging`` block
        (level, file, use_console, backup_count, max_size [KB], format, datefmt)."""
        log_config = self.config.get("logging", {}) if hasattr(self.config, "get") else {}
        level = str(log_config.get("level", "INFO")).upper()
        level = getattr(logging, level, logging.INFO) if level != "NOTSET" else logging.NOTSET
        fmt = log_config.get("format", "%(asctime)s - %(name)-14s - %(levelname)-8s - %(message)s")
        datefmt = log_config.get("datefmt", "%Y-%m-%d %H:%M:%S")
        self.log_dir.mkdir(parents=True, exist_ok=True)
        root = logging.getLogger()
        root.setLevel(level)
        for h in list(root.handlers):
            if getattr(h, "_v6b200", False):
                root.removeHandler(h)
                h.close()
        rfh = logging.handlers.RotatingFileHandler(
            str(self.log_file), maxBytes=1024 * int(log_config.get("max_size", 1024)),
            backupCount=int(log_config.get("backup_count", 5)))
        rfh.setLevel(level)
        rfh.setFormatter(logging.Formatter(fmt, datefmt))
        rfh._v6b200 = True  # type: ignore[attr-defined]
        root.addHandler(rfh)
        if log_config.get("use_console", False):
            ch = logging.StreamHandler(sys.stdout)
            ch.setLevel(level)
            ch.setFormatter(logging.Formatter(fmt, datefmt))
            ch._v6b200 = True  # type: ignore[attr-defined]
            root.addHandler(ch)
        self.log = logging.getLogger(logger_name(__name__))
        self.log.info("#" * 80)
        self.log.info(f"#{APPNAME:^78}#")
        self.log.info("#" * 80)
        self.log.info(f"Started application {APPNAME} with environment {self.environment}")
        self.log.info(f"Current working directory is '{os.getcwd()}'")
        self.log.info(f"Successfully loaded configuration from '{self.config_file}'")
        self.log.info(f"Logging to '{self.log_file}'")
