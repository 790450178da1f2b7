"""Entity store of the central server (sqlite3, stdlib only).

Entities and relations follow vantage6 3.x: Organization (public key),
Collaboration (set of organizations, ``encrypted`` flag), Node (one per organization x
collaboration, authenticates with an api_key), User (+ roles / rules), Task (name, image,
collaboration, run_id, parent, database label, initiator) and Result (one per task x
organization: input, result, log, assigned/started/finished timestamps).  "A task is complete
when all its results are finished."

The database file is the persistent state that survives restarts
(reference vantage6/cli/server.py:177-201, vantage6/cli/context.py:30-42).
The reference reaches it through SQLAlchemy URIs; ``sqlite:///rel`` and ``sqlite:////abs`` are
understood here, other engines are rejected with a clear message.
"""
from __future__ import annotations

import datetime as _dt
import hashlib
import hmac
import os
import secrets
import sqlite3
import threading
import uuid
from pathlib import Path
from typing import Any, Dict, Iterable, List, Optional

SCHEMA = """
CREATE TABLE IF NOT EXISTS organization (
  id INTEGER PRIMARY KEY AUTOINCREMENT, name TEXT UNIQUE, domain TEXT, address1 TEXT, address2 TEXT,
  zipcode TEXT, country TEXT, public_key TEXT);
CREATE TABLE IF NOT EXISTS collaboration (
  id INTEGER PRIMARY KEY AUTOINCREMENT, name TEXT UNIQUE, encrypted INTEGER DEFAULT 0);
CREATE TABLE IF NOT EXISTS member (
  collaboration_id INTEGER, organization_id INTEGER, PRIMARY KEY (collaboration_id, organization_id));
CREATE TABLE IF NOT EXISTS node (
  id INTEGER PRIMARY KEY AUTOINCREMENT, name TEXT, api_key TEXT UNIQUE, collaboration_id INTEGER,
  organization_id INTEGER, ip TEXT, status TEXT DEFAULT 'offline', last_seen TEXT, gpu INTEGER);
CREATE TABLE IF NOT EXISTS user (
  id INTEGER PRIMARY KEY AUTOINCREMENT, username TEXT UNIQUE, password TEXT, firstname TEXT, lastname TEXT,
  email TEXT, organization_id INTEGER, last_seen TEXT, failed_login_attempts INTEGER DEFAULT 0);
CREATE TABLE IF NOT EXISTS role (
  id INTEGER PRIMARY KEY AUTOINCREMENT, name TEXT, description TEXT, organization_id INTEGER);
CREATE TABLE IF NOT EXISTS rule (
  id INTEGER PRIMARY KEY AUTOINCREMENT, name TEXT, operation TEXT, scope TEXT, description TEXT,
  UNIQUE (name, operation, scope));
CREATE TABLE IF NOT EXISTS role_rule (role_id INTEGER, rule_id INTEGER, PRIMARY KEY (role_id, rule_id));
CREATE TABLE IF NOT EXISTS user_role (user_id INTEGER, role_id INTEGER, PRIMARY KEY (user_id, role_id));
CREATE TABLE IF NOT EXISTS user_rule (user_id INTEGER, rule_id INTEGER, PRIMARY KEY (user_id, rule_id));
CREATE TABLE IF NOT EXISTS task (
  id INTEGER PRIMARY KEY AUTOINCREMENT, name TEXT, description TEXT, image TEXT, collaboration_id INTEGER,
  run_id INTEGER, parent_id INTEGER, database TEXT, initiator_id INTEGER, init_user_id INTEGER,
  created_at TEXT);
CREATE TABLE IF NOT EXISTS result (
  id INTEGER PRIMARY KEY AUTOINCREMENT, task_id INTEGER, organization_id INTEGER, input TEXT, result TEXT,
  log TEXT, assigned_at TEXT, started_at TEXT, finished_at TEXT, status TEXT DEFAULT 'pending');
CREATE TABLE IF NOT EXISTS setting (key TEXT PRIMARY KEY, value TEXT);
CREATE TABLE IF NOT EXISTS port (
  id INTEGER PRIMARY KEY AUTOINCREMENT, result_id INTEGER, port INTEGER, label TEXT, address TEXT);
CREATE INDEX IF NOT EXISTS idx_result_task ON result(task_id);
CREATE INDEX IF NOT EXISTS idx_result_org ON result(organization_id);
"""

TABLES = ["organization", "collaboration", "member", "node", "user", "role", "rule", "role_rule", "user_role",
          "user_rule", "task", "result", "port"]


def now() -> str:
    return _dt.datetime.now(_dt.timezone.utc).isoformat()


def hash_password(password: str) -> str:
    salt = os.urandom(16)
    dk = hashlib.pbkdf2_hmac("sha256", password.encode("utf-8"), salt, 60_000)
    return f"pbkdf2${salt.hex()}${dk.hex()}"


def check_password(password: str, stored: str) -> bool:
    try:
        _, salt, dk = stored.split("$")
        test = hashlib.pbkdf2_hmac("sha256", password.encode("utf-8"), bytes.fromhex(salt), 60_000)
        return hmac.compare_digest(test.hex(), dk)
    except Exception:  # noqa: BLE001
        return False


def sqlite_path_from_uri(uri: str, data_dir: Optional[Path] = None) -> str:
    if not uri.startswith("sqlite"):
        raise ValueError(f"unsupported database URI {uri!r}: this server stores its entities in sqlite "
                         "(use sqlite:///relative.sqlite or sqlite:////absolute/path.sqlite)")
    if ":///" not in uri:
        return ":memory:"
    path = uri.split(":///", 1)[1]
    if not path:
        return ":memory:"
    if not os.path.isabs(path) and data_dir is not None:
        path = str(Path(data_dir) / path)
    return path


class Database:
    """Thread-safe sqlite wrapper returning plain dict rows."""

    def __init__(self, uri: str = "sqlite://", data_dir: Optional[Path] = None, allow_drop_all: bool = False):
        self.path = sqlite_path_from_uri(uri, data_dir)
        if self.path != ":memory:":
            Path(self.path).parent.mkdir(parents=True, exist_ok=True)
        self.allow_drop_all = allow_drop_all
        self._lock = threading.RLock()
        self._conn = sqlite3.connect(self.path, check_same_thread=False, isolation_level=None)
        self._conn.row_factory = sqlite3.Row
        self._conn.execute("PRAGMA journal_mode=WAL")
        self._conn.execute("PRAGMA synchronous=NORMAL")
        with self._lock:
            self._conn.executescript(SCHEMA)
            self._migrate()

    # columns added after the first release: an existing database file gains them on open
    MIGRATIONS = {"user": {"last_login_attempt": "TEXT"}}

    def _migrate(self) -> None:
        for table, columns in self.MIGRATIONS.items():
            have = {r["name"] for r in self._conn.execute(f"PRAGMA table_info({table})").fetchall()}
            for col, decl in columns.items():
                if col not in have:
                    self._conn.execute(f"ALTER TABLE {table} ADD COLUMN {col} {decl}")

    # -- primitives --------------------------------------------------------------------------
    def execute(self, sql: str, args: Iterable = ()) -> sqlite3.Cursor:
        with self._lock:
            return self._conn.execute(sql, tuple(args))

    def query(self, sql: str, args: Iterable = ()) -> List[Dict[str, Any]]:
        with self._lock:
            return [dict(r) for r in self._conn.execute(sql, tuple(args)).fetchall()]

    def one(self, sql: str, args: Iterable = ()) -> Optional[Dict[str, Any]]:
        rows = self.query(sql, args)
        return rows[0] if rows else None

    @staticmethod
    def _columns(table: str, fields: Dict[str, Any]) -> List[str]:
        """Table and column names are interpolated into SQL (values never are): only known tables and plain
        identifiers get through."""
        if table not in TABLES or not all(isinstance(k, str) and k.isidentifier() for k in fields):
            raise ValueError(f"refusing SQL identifiers {table!r} / {sorted(fields)!r}")
        return list(fields)

    def insert(self, table: str, **fields) -> int:
        keys = self._columns(table, fields)
        sql = f"INSERT INTO {table} ({','.join(keys)}) VALUES ({','.join('?' * len(keys))})"
        with self._lock:
            cur = self._conn.execute(sql, [fields[k] for k in keys])
            return int(cur.lastrowid)

    def update(self, table: str, id_: int, **fields) -> None:
        if not self._columns(table, fields):
            return
        sets = ",".join(f"{k}=?" for k in fields)
        self.execute(f"UPDATE {table} SET {sets} WHERE id=?", list(fields.values()) + [id_])

    def get(self, table: str, id_: int) -> Optional[Dict[str, Any]]:
        assert table in TABLES
        return self.one(f"SELECT * FROM {table} WHERE id=?", (id_,))

    def delete(self, table: str, id_: int) -> None:
        assert table in TABLES
        self.execute(f"DELETE FROM {table} WHERE id=?", (id_,))

    def drop_all(self) -> None:
        if not self.allow_drop_all:
            raise PermissionError("allow_drop_all is False in the server configuration")
        with self._lock:
            for t in TABLES:
                self._conn.execute(f"DROP TABLE IF EXISTS {t}")
            self._conn.executescript(SCHEMA)
            self._migrate()

    def close(self) -> None:
        with self._lock:
            self._conn.close()

    # -- domain helpers ----------------------------------------------------------------------
    def collaboration_organizations(self, collaboration_id: int) -> List[int]:
        return [r["organization_id"] for r in
                self.query("SELECT organization_id FROM member WHERE collaboration_id=? ORDER BY organization_id",
                           (collaboration_id,))]

    def organization_collaborations(self, organization_id: int) -> List[int]:
        return [r["collaboration_id"] for r in
                self.query("SELECT collaboration_id FROM member WHERE organization_id=? ORDER BY collaboration_id",
                           (organization_id,))]

    def next_run_id(self) -> int:
        r = self.one("SELECT MAX(run_id) AS m FROM task")
        return int(r["m"] or 0) + 1

    def new_api_key(self) -> str:
        return str(uuid.uuid1())

    def task_complete(self, task_id: int) -> bool:
        r = self.one("SELECT COUNT(*) AS n FROM result WHERE task_id=? AND finished_at IS NULL", (task_id,))
        return int(r["n"]) == 0

    def user_rules(self, user_id: int) -> List[Dict[str, Any]]:
        return self.query(
            "SELECT DISTINCT rule.* FROM rule WHERE rule.id IN (SELECT rule_id FROM user_rule WHERE user_id=?) "
            "OR rule.id IN (SELECT rule_id FROM role_rule WHERE role_id IN (SELECT role_id FROM user_role WHERE user_id=?))",
            (user_id, user_id))

    def token_secret(self, configured: Optional[str]) -> str:
        """JWT secret: the configured constant (``jwt_secret_key``) or a random one kept with the database, so that
        tokens outlive a server restart and a second process on the same database (``vserver shell``, a second server
        behind the message queue) signs and verifies with the same key.  ``drop_all`` does not touch it."""
        if configured:
            return configured
        with self._lock:
            self._conn.execute("INSERT OR IGNORE INTO setting VALUES ('jwt_secret_key', ?)", (secrets.token_hex(32),))
            return self._conn.execute("SELECT value FROM setting WHERE key='jwt_secret_key'").fetchone()[0]
