"""Multi-GPU tier: collected by ``pytest -m gpu`` and skipped on boxes with one GPU.

* the NVLink data-plane correctness suite (tests/dist_comm_check.py: heap bring-up, P2P / multicast copies, K2 for
  every server optimizer / upload mode / placement with unequal weights and a non-reporting node, K3, K1),
* the failure path on real flags (tests/dist_fault_check.py: a rank dies mid-round -> in-kernel timeout -> status and
  missing set reach every survivor -> mark_dead + re-sharding -> the round completes with renormalised weights).
"""
import glob
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NGPU = torch.cuda.device_count() if torch.cuda.is_available() else 0
needs2 = pytest.mark.skipif(NGPU < 2, reason="needs >= 2 GPUs")


def _torchrun(script, nproc, port, *args, timeout=600):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, script), *args]
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    return subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)


@needs2
def test_nvlink_data_plane_suite(tmp_path):
    n = min(NGPU, 8)
    out = tmp_path / "comm.json"
    pr = _torchrun("tests/dist_comm_check.py", n, 29541, "--out", str(out), "--big", "2000000")
    assert pr.returncode == 0, pr.stdout[-3000:]
    res = json.loads(out.read_text().strip().splitlines()[-1])
    assert res["world"] == n and res["k2_checks_passed"] == 48 and res["k3_us"] > 0, res


@needs2
@pytest.mark.parametrize("mode,upload", [("sharded", "weights_f32"), ("central", "delta_f32")])
def test_rank_killed_mid_round_is_excluded_and_round_completes(tmp_path, mode, upload):
    n = min(NGPU, 4)
    base = str(tmp_path / "fault")
    pr = _torchrun("tests/dist_fault_check.py", n, 29543 if mode == "sharded" else 29545, "--out", base, "--server-mode", mode,
                   "--upload", upload, timeout=300)
    logs = [json.load(open(f)) for f in sorted(glob.glob(base + "_rank*.json"))]
    assert len(logs) == n - 1, pr.stdout[-3000:]                       # every survivor finished
    for lg in logs:
        assert lg["r1_status"] == 0 and abs(lg["r1"][0] - lg["r1_expected"]) < 1e-3 and abs(lg["r1"][1] - lg["r1_expected"]) < 1e-3, lg
        assert lg["r2_status"] != 0 and (lg["r2_missing"] >> (n - 1)) & 1, lg     # the timeout was seen, the victim identified
        assert lg["newly_dead"] == [n - 1] and lg["dead"] == [n - 1], lg
        assert abs(lg["r2_recovered"][0] - lg["r2_expected"]) < 1e-3 and abs(lg["r2_recovered"][1] - lg["r2_expected"]) < 1e-3, lg
        assert lg["r3_status"] == 0 and abs(lg["r3"][0] - lg["r3_expected"]) < 1e-3 and abs(lg["r3"][1] - lg["r3_expected"]) < 1e-3, lg
        assert (n - 1) not in lg["reducers"], lg
