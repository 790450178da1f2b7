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
           return_weights: bool = False, checkpoint_every: int = 0, checkpoint_dir: Optional[str] = None,
           resume_from: Optional[str] = None, fault_tolerant: bool = False, timeout_ms: Optional[float] = None,
           data_source: str = "auto") -> Dict[str, Any]:
    orgs = client.get_organizations_in_my_collaboration()
    ids = sorted(organization_ids or [o["id"] for o in orgs])
    rendezvous = {"addr": "127.0.0.1", "port": _free_port(), "world": len(ids), "ranks": {str(o): r for r, o in enumerate(ids)}}
    task = client.create_new_task(
        input_={"method": "train", "kwargs": dict(model=model, rounds=rounds, local_steps=local_steps, batch=batch,
                                                  server_opt=server_opt, server_lr=server_lr, server_mode=server_mode,
                                                  rendezvous=rendezvous, seed=seed, return_weights=return_weights,
                                                  checkpoint_every=checkpoint_every, checkpoint_dir=checkpoint_dir,
                                                  resume_from=resume_from, fault_tolerant=fault_tolerant, timeout_ms=timeout_ms,
                                                  data_source=data_source)},
        organization_ids=ids, name=f"fedavg-{model}")
    client.wait_for_task(task["id"])
    results = client.get_results(task_id=task["id"])
    results = sorted((r for r in results if r), key=lambda r: r["rank"])
    n_r = min((len(r["losses"]) for r in results), default=0)
    losses = [sum(r["losses"][i] * r["n_samples"] for r in results) / sum(r["n_samples"] for r in results)
              for i in range(n_r)] if results else []
    out = {"model": model, "rounds": rounds, "world": len(ids), "global_loss": losses,
           "rounds_per_sec": min(r["rounds_per_sec"] for r in results) if results else None,
           "ms_per_round_max": max(r["ms_per_round"] for r in results) if results else None,
           "data_plane": results[0]["data_plane"] if results else None,
           "multicast": results[0].get("multicast") if results else None,
           "nodes": [{k: r.get(k) for k in ("rank", "organization_id", "n_samples", "device", "data_source", "trainer_reused",
                                            "setup_s")} for r in results]}
    if return_weights and results:
        out["weights_checksum"] = [r.get("weights_checksum") for r in results]
    return out


def _confined(path: Optional[str], what: str) -> Optional[str]:
    """Task inputs come from researchers: a path they name (checkpoints, metrics) may only point into directories the node
    set aside for algorithms -- its per-run temporary folder, its log directory, or ``V6_ALGORITHM_DATA_DIR`` (a colon-separated
    list the node operator adds).  Relative paths are taken inside the first of them.  Anything else is refused (the reference
    confines algorithms to container mounts: reference vantage6/cli/node.py:360-378)."""
    if not path:
        return None
    roots = [os.path.realpath(r) for r in (os.environ.get("TEMPORARY_FOLDER"), os.environ.get("V6_LOG_DIR"),
                                           *(os.environ.get("V6_ALGORITHM_DATA_DIR", "").split(os.pathsep))) if r]
    if not roots:
        raise PermissionError(f"{what}: the node exported no directory an algorithm may write to")
    p = path if os.path.isabs(path) else os.path.join(roots[0], path)
    real = os.path.realpath(p)
    if not any(real == r or real.startswith(r + os.sep) for r in roots):
        raise PermissionError(f"{what}={path!r} is outside the directories this node allows algorithms to use")
    return real


def RPC_train(data, **kwargs) -> Dict[str, Any]:
    """Node-side partial.  On a node with a resident GPU worker (``vnode start --gpu K``, node/gpu_worker.py) the
    request is forwarded to it -- CUDA context, symmetric heap, model and CUDA graphs live there across tasks --
    otherwise the training runs in this process."""
    sock = os.environ.get("V6_GPU_WORKER")
    if sock and os.path.exists(sock):
        from ...node.gpu_worker import call

        label = os.environ.get("DATABASE_LABEL", "default").upper()
        uri = os.environ.get(f"{label}_DATABASE_URI") or os.environ.get("DATABASE_URI")
        return call(sock, {"op": "train", "kwargs": kwargs, "organization_id": int(os.environ.get("V6_ORGANIZATION_ID", "0")),
                           "database_uri": uri})
    return train_partial(data, **kwargs)


def train_partial(data, model: str = "resnet_tiny", rounds: int = 2, local_steps: Optional[int] = None,
                  batch: Optional[int] = None, server_opt: str = "fedavg", server_lr: float = 1.0,
                  server_mode: str = "sharded", rendezvous: Optional[dict] = None, seed: int = 0,
                  return_weights: bool = False, checkpoint_every: int = 0, checkpoint_dir: Optional[str] = None,
                  resume_from: Optional[str] = None, fault_tolerant: bool = False, timeout_ms: Optional[float] = None,
                  metrics_file: Optional[str] = None, data_source: str = "auto", trainer_cache=None) -> Dict[str, Any]:
    """All R rounds of one node, device-resident.  ``data`` is the node's labelled database (algorithm/data.py; a
    ``synthetic://`` URI or no database falls back to the model's synthetic generator, seeded per organization).
    ``checkpoint_every`` / ``checkpoint_dir`` / ``resume_from`` drive utils/checkpoint.py; per-round metrics rows go to
    ``metrics_file`` (default: ``$V6_LOG_DIR/fedavg_metrics.jsonl`` when the node exports its log directory)."""
    import torch
    import torch.distributed as dist

    checkpoint_dir = _confined(checkpoint_dir, "checkpoint_dir")
    resume_from = _confined(resume_from, "resume_from")
    metrics_file = _confined(metrics_file, "metrics_file")
    from ...models import zoo
    from ...parallel.fedavg import ServerOptConfig
    from ...utils.metrics import MetricsWriter
    from ..data import make_local_batches

    t_start = time.perf_counter()
    org_id = int(os.environ.get("V6_ORGANIZATION_ID", "0"))
    rv = rendezvous or {"addr": "127.0.0.1", "port": _free_port(), "world": 1, "ranks": {str(org_id): 0}}
    rank, world = int(rv["ranks"][str(org_id)]), int(rv["world"])
    use_cuda = torch.cuda.is_available()
    device = torch.device("cuda", int(os.environ.get("V6_GPU", "0"))) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(device)
    key = (model, world, rank, server_mode, server_opt, float(server_lr), tuple(sorted((str(k), int(v)) for k, v in rv["ranks"].items())))
    cached = trainer_cache.get_trainer(key) if trainer_cache is not None else None
    created_pg = False
    reused = cached is not None
    if cached is not None:
        tr, spec = cached
        tr.reset(seed)
    else:
        # the native data plane needs no process group at all (the symmetric heap has its own rendezvous); the CPU /
        # collective plane runs on gloo
        if world > 1 and not use_cuda and not dist.is_initialized():
            dist.init_process_group("gloo", init_method=f"tcp://{rv['addr']}:{rv['port']}", rank=rank, world_size=world)
            created_pg = True
        os.environ["MASTER_PORT"] = str(rv["port"])      # symmetric-heap rendezvous directory key
        torch.manual_seed(seed)
        extra = {}
        if timeout_ms is not None:
            extra["timeout_ms"] = float(timeout_ms)
        tr, spec = zoo.build_trainer(model, rank=rank, world=world, device=device, server_mode=server_mode,
                                     server_opt=ServerOptConfig(server_opt, server_lr), fault_tolerant=fault_tolerant, **extra)
        tr.remember_init(seed)
        if trainer_cache is not None:
            trainer_cache.put_trainer(key, (tr, spec))
    log_dir = os.environ.get("V6_LOG_DIR")
    if metrics_file is None and log_dir:
        metrics_file = os.path.join(log_dir, "fedavg_metrics.jsonl")
    tr.metrics = MetricsWriter(metrics_file, static={"model": model, "organization_id": org_id}) if metrics_file else None
    tr.checkpoint_dir, tr.checkpoint_every = checkpoint_dir, int(checkpoint_every or 0)
    tr.fault_tolerant = bool(fault_tolerant)
    bsz = batch or spec.batch
    batches, n_samples, source = make_local_batches(data, spec, local_steps, bsz, 1000 + org_id, pin=use_cuda, source=data_source)
    first_round = 0
    if resume_from:
        first_round = tr.load_checkpoint(resume_from)       # the next aggregation re-broadcasts the global model
    else:
        tr.initialize_global()
    setup_s = time.perf_counter() - t_start
    losses = []
    t0 = None
    for r in range(first_round, rounds):
        if r == first_round + 1 or rounds - first_round == 1:
            if use_cuda:
                torch.cuda.synchronize()
            t0 = time.perf_counter()
        t_r = time.perf_counter()
        losses.append(float(tr.run_round(batches, n_samples).item()))
        if tr.metrics is not None:
            tr.metrics.round(r + 1, 1e3 * (time.perf_counter() - t_r), losses[-1], tr.engine.nvlink_bytes_per_round(), rank=rank)
    if use_cuda:
        torch.cuda.synchronize()
    done = rounds - first_round
    timed_rounds = max(1, done - 1) if done > 1 else 1
    dt = (time.perf_counter() - t0) if t0 is not None else float("nan")
    if tr.fault_tolerant:
        tr.recover_if_failed()
    out = {"rank": rank, "organization_id": org_id, "losses": losses, "n_samples": n_samples, "device": str(device),
           "rounds_per_sec": timed_rounds / dt, "ms_per_round": 1e3 * dt / timed_rounds,
           "data_plane": tr.engine.data_plane, "multicast": bool(tr.engine.use_multicast),
           "comm_status": tr.engine.poll_status(), "data_source": source, "trainer_reused": reused,
           "setup_s": setup_s, "first_round": first_round, "dead": list(tr.dead)}
    if return_weights:
        out["weights_checksum"] = float(tr.engine.w.double().sum().item())
    if trainer_cache is None:
        tr.close()
        if created_pg:
            dist.barrier()
            dist.destroy_process_group()
    return out
