"""FedAvg as a vantage6 algorithm (BASELINE configs 2-4).

Control plane (vantage6 semantics): the researcher creates ONE task
``{"method": "master", "master": true, "kwargs": {"model": "resnet50", "rounds": R, ...}}``;
the ``master`` (running on one node) creates a ``train`` sub-task for every organization and
waits for their results -- task dispatch, bookkeeping and result collection go through the
server exactly like any vantage6 algorithm.

Data plane (B200): the ``train`` partials rendezvous (rank = position of the organization in
the sorted participant list) and run all R rounds device-resident: local steps under a CUDA
graph, then ONE fused aggregation kernel per round over NVLink symmetric memory.  Node data
never leaves its GPU; only model parameters / deltas cross, and they never touch the REST
plane.  On CPU nodes the same code runs on gloo with the collective data plane.
"""
from __future__ import annotations

import os
import socket
import time
from typing import Any, Dict, List, Optional


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def master(client, data, model: str = "resnet_tiny", rounds: int = 2, local_steps: Optional[int] = None,
           batch: Optional[int] = None, server_opt: str = "fedavg", server_lr: float = 1.0,
           server_mode: str = "sharded", organization_ids: Optional[List[int]] = None, seed: int = 0,
           return_weights: bool = False) -> Dict[str, Any]:
    orgs = client.get_organizations_in_my_collaboration()
    ids = sorted(organization_ids or [o["id"] for o in orgs])
    rendezvous = {"addr": "127.0.0.1", "port": _free_port(), "world": len(ids), "ranks": {str(o): r for r, o in enumerate(ids)}}
    task = client.create_new_task(
        input_={"method": "train", "kwargs": dict(model=model, rounds=rounds, local_steps=local_steps, batch=batch,
                                                  server_opt=server_opt, server_lr=server_lr, server_mode=server_mode,
                                                  rendezvous=rendezvous, seed=seed, return_weights=return_weights)},
        organization_ids=ids, name=f"fedavg-{model}")
    client.wait_for_task(task["id"])
    results = client.get_results(task_id=task["id"])
    results = sorted((r for r in results if r), key=lambda r: r["rank"])
    losses = [sum(r["losses"][i] * r["n_samples"] for r in results) / sum(r["n_samples"] for r in results)
              for i in range(rounds)] if results else []
    out = {"model": model, "rounds": rounds, "world": len(ids), "global_loss": losses,
           "rounds_per_sec": min(r["rounds_per_sec"] for r in results) if results else None,
           "ms_per_round_max": max(r["ms_per_round"] for r in results) if results else None,
           "data_plane": results[0]["data_plane"] if results else None,
           "multicast": results[0].get("multicast") if results else None,
           "nodes": [{k: r[k] for k in ("rank", "organization_id", "n_samples", "device")} for r in results]}
    if return_weights and results:
        out["weights_checksum"] = [r.get("weights_checksum") for r in results]
    return out


def RPC_train(data, model: str = "resnet_tiny", rounds: int = 2, local_steps: Optional[int] = None,
              batch: Optional[int] = None, server_opt: str = "fedavg", server_lr: float = 1.0,
              server_mode: str = "sharded", rendezvous: Optional[dict] = None, seed: int = 0,
              return_weights: bool = False) -> Dict[str, Any]:
    import torch
    import torch.distributed as dist

    from ...models import zoo
    from ...parallel.fedavg import ServerOptConfig

    org_id = int(os.environ.get("V6_ORGANIZATION_ID", "0"))
    rv = rendezvous or {"addr": "127.0.0.1", "port": _free_port(), "world": 1, "ranks": {str(org_id): 0}}
    rank, world = int(rv["ranks"][str(org_id)]), int(rv["world"])
    use_cuda = torch.cuda.is_available()
    device = torch.device("cuda", int(os.environ.get("V6_GPU", "0"))) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(device)
    created_pg = False
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl" if use_cuda else "gloo", init_method=f"tcp://{rv['addr']}:{rv['port']}",
                                rank=rank, world_size=world, **({"device_id": device} if use_cuda else {}))
        created_pg = True
        os.environ["MASTER_PORT"] = str(rv["port"])      # symmetric-heap rendezvous directory key
    torch.manual_seed(seed)
    tr, spec = zoo.build_trainer(model, rank=rank, world=world, device=device, server_mode=server_mode,
                                 server_opt=ServerOptConfig(server_opt, server_lr))
    n_steps = local_steps or spec.local_steps
    bsz = batch or spec.batch
    # the node's own data: a synthetic shard seeded by the organization (non-IID across nodes)
    batches = spec.make_batches(n_steps, bsz, 1000 + org_id, pin=use_cuda)
    n_samples = float(n_steps * bsz)
    tr.initialize_global()
    losses = []
    t0 = None
    for r in range(rounds):
        if r == 1 or rounds == 1:
            if use_cuda:
                torch.cuda.synchronize()
            t0 = time.perf_counter()
        losses.append(float(tr.run_round(batches, n_samples).item()))
    if use_cuda:
        torch.cuda.synchronize()
    timed_rounds = max(1, rounds - 1) if rounds > 1 else 1
    dt = (time.perf_counter() - t0) if t0 is not None else float("nan")
    out = {"rank": rank, "organization_id": org_id, "losses": losses, "n_samples": n_samples, "device": str(device),
           "rounds_per_sec": timed_rounds / dt, "ms_per_round": 1e3 * dt / timed_rounds,
           "data_plane": tr.engine.data_plane, "multicast": bool(tr.engine.use_multicast),
           "comm_status": tr.engine.poll_status()}
    if return_weights:
        out["weights_checksum"] = float(tr.engine.w.double().sum().item())
    tr.close()
    if created_pg:
        dist.barrier()
        dist.destroy_process_group()
    return out
