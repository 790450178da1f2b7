"""The warm-start helper of the node runtime (node/zygote.py): forked algorithm processes behave like the
fresh-interpreter ones (environment contract, output file, exit codes, log capture, kill / timeout)."""
import subprocess
import textwrap
import time

import pytest

from vantage6_b200.common.serialization import deserialize, serialize
from vantage6_b200.node.zygote import Zygote


@pytest.fixture()
def zygote(tmp_path):
    (tmp_path / "zyg_algo.py").write_text(textwrap.dedent("""
        import os, time

        def RPC_echo(data, x):
            print("hello from", os.getpid())
            return {"x": x, "org": os.environ.get("V6_ORGANIZATION_ID"), "data": data}

        def RPC_sleep(data, seconds):
            time.sleep(seconds)
            return "done"

        def RPC_boom(data):
            raise ValueError("algorithm failure")
    """))
    z = Zygote(tmp_path)
    assert z.start(), "zygote did not come up"
    yield z, tmp_path
    z.stop()


def _task(tmp_path, name, method, **kwargs):
    work = tmp_path / name
    work.mkdir()
    (work / "input").write_bytes(serialize({"method": method, "kwargs": kwargs}, "json"))
    (work / "token").write_text("t")
    env = {"INPUT_FILE": str(work / "input"), "OUTPUT_FILE": str(work / "output"), "TOKEN_FILE": str(work / "token"),
           "PYTHONPATH": str(tmp_path), "V6_ORGANIZATION_ID": "42", "DATABASE_URI": "synthetic://x", "PATH": "/usr/bin:/bin"}
    return work, env


def test_forked_algorithm_honours_the_environment_contract(zygote):
    z, tmp = zygote
    work, env = _task(tmp, "a", "echo", x=3)
    t0 = time.time()
    proc = z.spawn("zyg_algo", env, work / "log")
    assert proc.pid > 0 and proc.wait(timeout=30) == 0
    assert time.time() - t0 < 5.0
    assert deserialize((work / "output").read_bytes()) == {"x": 3, "org": "42", "data": "synthetic://x"}
    assert "hello from" in proc.read_log()


def test_failures_and_timeouts_are_reported(zygote):
    z, tmp = zygote
    work, env = _task(tmp, "b", "boom")
    proc = z.spawn("zyg_algo", env, work / "log")
    assert proc.wait(timeout=30) == 1
    assert "algorithm failure" in proc.read_log()
    work, env = _task(tmp, "c", "sleep", seconds=30)
    proc = z.spawn("zyg_algo", env, work / "log")
    with pytest.raises(subprocess.TimeoutExpired):
        proc.wait(timeout=0.3)
    proc.kill()
    assert proc.wait(timeout=10) != 0


def test_children_run_concurrently(zygote):
    z, tmp = zygote
    procs = []
    for i in range(4):
        work, env = _task(tmp, f"p{i}", "sleep", seconds=0.5)
        procs.append(z.spawn("zyg_algo", env, work / "log"))
    t0 = time.time()
    assert [p.wait(timeout=30) for p in procs] == [0, 0, 0, 0]
    assert time.time() - t0 < 1.8          # four 0.5 s sleeps side by side, not back to back


def test_preload_follows_the_nodes_databases():
    from vantage6_b200.node.zygote import PRELOAD, preload_for

    assert "pandas" not in preload_for(["/data/vec.npy", "synthetic://imagenet", "/data/shard.pt"])
    assert "pandas" in preload_for(["/data/vec.npy", "/data/patients.CSV"])
    assert "pandas" in preload_for(["/data/x.parquet"])
    assert preload_for([]) == PRELOAD and preload_for(None) == PRELOAD
