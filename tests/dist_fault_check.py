#!/usr/bin/env python
"""Failure path on real NVLink flags (run under torchrun, >= 2 GPUs; no NCCL process group is created):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tests/dist_fault_check.py --out gpurun_out/fault_N

round 1: every node reports (unequal sample counts n_i, each rank only knows its own) -> sample-weighted mean
round 2: the LAST rank dies after its local steps, before the aggregation.  The survivors' fused aggregation kernel
         times out on its upload flag (short timeout), publishes status + the missing rank to every peer, nobody pushes;
         at the start of round 3 `FederatedTrainer.recover_if_failed` marks the rank dead on every survivor, re-partitions
         the slices and re-runs round 2's aggregation over the survivors (renormalised weights)
round 3: a normal round among the survivors.
Each survivor writes {rank, dead, w_after_round_k, expected_k} to <out>_rank<r>.json; the pytest parent compares.
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/fault")
    ap.add_argument("--server-mode", default="sharded")
    ap.add_argument("--upload", default="weights_f32")
    ap.add_argument("--timeout-ms", type=float, default=1500.0)
    a = ap.parse_args()
    rank, world, lr_ = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr_)
    dev = torch.device("cuda", lr_)
    from torch import nn

    from vantage6_b200.parallel.trainer import FederatedTrainer

    torch.manual_seed(0)
    model = nn.Linear(64, 64, bias=False)                      # 4096 parameters: several slices even at 8 ranks
    tr = FederatedTrainer(model, lambda m, x, y: ((m(x) - y) ** 2).mean(), rank=rank, world=world, device=dev, optimizer="sgd",
                          lr=0.0, momentum=0.0, server_mode=a.server_mode, upload=a.upload, amp_dtype=None,
                          use_cuda_graph=False, fault_tolerant=True, timeout_ms=a.timeout_ms)
    tr.initialize_global()
    x, y = torch.zeros(2, 64, device=dev), torch.zeros(2, 64, device=dev)
    n_i = 100.0 * (rank + 1)
    nt = tr.fm.n_trainable
    victim = world - 1
    log = {"rank": rank, "world": world, "mode": a.server_mode, "upload": a.upload}

    def local_value(rnd):                                      # what node `rank` "trains" its weights to in round rnd
        return float(rnd * 10 + rank + 1)

    orig = tr.engine.aggregate
    state = {"rnd": 1}

    def agg(*args, **kw):                                      # lr = 0: set the local result by hand right before aggregating
        with torch.no_grad():
            v = local_value(state["rnd"])
            if a.upload == "weights_f32":
                tr.fm.params.fill_(v)
            else:
                tr.engine.upload[:nt].copy_(((v - tr.w_ref[:nt]) * n_i).to(tr.engine.upload.dtype))
        return orig(*args, **kw)
    tr.engine.aggregate = agg

    def expected(rnd, ranks):
        return sum(100.0 * (r + 1) * (rnd * 10 + r + 1) for r in ranks) / sum(100.0 * (r + 1) for r in ranks)

    # ---- round 1: everybody
    tr.run_round([(x, y)], n_samples=n_i)
    torch.cuda.synchronize()
    log["r1"] = [float(tr.fm.params.min()), float(tr.fm.params.max())]
    log["r1_expected"] = expected(1, range(world))
    log["r1_status"] = tr.engine.poll_status()
    # ---- round 2: the victim dies before the aggregation
    state["rnd"] = 2
    if rank == victim:
        time.sleep(0.2)
        os._exit(0)
    t0 = time.time()
    tr.run_round([(x, y)], n_samples=n_i)
    torch.cuda.synchronize()
    log["r2_wait_s"] = time.time() - t0
    log["r2_status"] = tr.engine.poll_status()
    log["r2_missing"] = tr.engine.missing_mask()
    # ---- round 3: recovery happens at the start (re-runs round 2's aggregation over the survivors), then a normal round
    newly = tr.recover_if_failed()
    torch.cuda.synchronize()
    log["newly_dead"] = newly
    log["r2_recovered"] = [float(tr.fm.params.min()), float(tr.fm.params.max())]
    log["r2_expected"] = expected(2, range(world - 1))
    state["rnd"] = 3
    tr.run_round([(x, y)], n_samples=n_i)
    torch.cuda.synchronize()
    log["r3"] = [float(tr.fm.params.min()), float(tr.fm.params.max())]
    log["r3_expected"] = expected(3, range(world - 1))
    log["r3_status"] = tr.engine.poll_status()
    log["dead"] = list(tr.dead)
    log["reducers"] = list(tr.engine.reducers)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(f"{a.out}_rank{rank}.json", "w") as f:
        json.dump(log, f)
    print(json.dumps(log), flush=True)
    os._exit(0)        # the victim's mappings are gone: skip the collective teardown


if __name__ == "__main__":
    main()
