"""The node's resident worker (node/gpu_worker.py) through its socket protocol, on the CPU data plane: a `train` request builds a
trainer, the next one for the same federation reuses it (the property the 8-GPU demo relies on: second task 0.57 s)."""
import os
import subprocess
import sys
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def worker(tmp_path):
    sock = str(tmp_path / "w.sock")
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""), CUDA_VISIBLE_DEVICES="")
    proc = subprocess.Popen([sys.executable, "-m", "vantage6_b200.node.gpu_worker", "--socket", sock], env=env,
                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    from vantage6_b200.node.gpu_worker import call

    deadline = time.time() + 120
    while time.time() < deadline:
        if proc.poll() is not None:
            raise RuntimeError("worker exited: " + proc.stdout.read().decode()[-2000:])
        if os.path.exists(sock):
            try:
                call(sock, {"op": "ping"}, timeout=5)
                break
            except Exception:  # noqa: BLE001
                pass
        time.sleep(0.2)
    else:
        proc.kill()
        raise RuntimeError("worker did not come up")
    yield sock, call
    try:
        call(sock, {"op": "shutdown"}, timeout=5)
        proc.wait(timeout=20)
    except Exception:  # noqa: BLE001
        proc.kill()


def test_worker_keeps_the_trainer_across_tasks(worker):
    sock, call = worker
    assert call(sock, {"op": "ping"}, timeout=5)["gpu"] is None
    kw = {"model": "resnet_tiny", "rounds": 1, "local_steps": 1, "batch": 4}
    a = call(sock, {"op": "train", "kwargs": kw, "organization_id": 1}, timeout=300)
    b = call(sock, {"op": "train", "kwargs": kw, "organization_id": 1}, timeout=300)
    assert a["trainer_reused"] is False and b["trainer_reused"] is True
    assert len(a["losses"]) == 1 and len(b["losses"]) == 1
    assert abs(a["losses"][0] - b["losses"][0]) < 1e-4          # reset(seed): the second task starts from the same model
    assert b["setup_s"] < a["setup_s"]
    st = call(sock, {"op": "stats"}, timeout=5)
    assert st["tasks"] == 2 and st["trainer_builds"] == 1 and st["trainer_reuses"] == 1


def test_worker_reports_errors_instead_of_dying(worker):
    sock, call = worker
    with pytest.raises(RuntimeError, match="unknown op"):
        call(sock, {"op": "nope"}, timeout=5)
    with pytest.raises(RuntimeError):
        call(sock, {"op": "train", "kwargs": {"model": "no_such_model"}, "organization_id": 1}, timeout=60)
    assert call(sock, {"op": "ping"}, timeout=5)["pid"] > 0      # still serving
