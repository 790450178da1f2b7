"""Federated logistic-regression GLM as a vantage6 algorithm (BASELINE config 5).

Two flavours:
* ``master`` / ``RPC_gradient``   -- classic vantage6 iteration on the CONTROL plane: every
  iteration is a sub-task round trip through the server (works on CPU nodes, any transport);
* ``master_fused`` / ``RPC_fit``  -- the B200 path: ONE sub-task; the partials rendezvous and
  iterate device-resident with K8 (fused gradient) + K3 (small all-reduce over NVLink).
"""
from __future__ import annotations

import os
import time
from typing import Any, Dict, Optional

import numpy as np

from .fedavg import _free_port


def _xy(data, features: Optional[int] = None):
    if isinstance(data, dict):
        return np.asarray(data["X"], dtype=np.float64), np.asarray(data["y"], dtype=np.float64)
    if hasattr(data, "to_numpy"):
        a = data.to_numpy(dtype=np.float64)
        return a[:, :-1], a[:, -1]
    a = np.asarray(data, dtype=np.float64)
    return a[:, :-1], a[:, -1]


# ------------------------------------------------------------------ control-plane flavour
def master(client, data, iterations: int = 10, lr: float = 1.0, organization_ids=None) -> Dict[str, Any]:
    ids = organization_ids or [o["id"] for o in client.get_organizations_in_my_collaboration()]
    w = None
    losses = []
    for _ in range(iterations):
        task = client.create_new_task(input_={"method": "gradient", "kwargs": {"w": None if w is None else w.tolist()}},
                                      organization_ids=ids)
        client.wait_for_task(task["id"])
        res = client.get_results(task_id=task["id"])
        g = sum(np.asarray(r["grad"], dtype=np.float64) for r in res)
        n = sum(r["n"] for r in res)
        if w is None:
            w = np.zeros_like(g)
        w = w - lr * g / n
        losses.append(sum(r["loss"] for r in res) / n)
    return {"coefficients": w[:-1], "intercept": float(w[-1]), "losses": losses, "n": int(n)}


def RPC_gradient(data, w=None) -> Dict[str, Any]:
    X, y = _xy(data)
    wv = np.zeros(X.shape[1] + 1) if w is None else np.asarray(w, dtype=np.float64)
    z = X @ wv[:-1] + wv[-1]
    p = 1.0 / (1.0 + np.exp(-z))
    r = p - y
    loss = float(np.sum(np.maximum(z, 0) - z * y + np.log1p(np.exp(-np.abs(z)))))
    return {"grad": np.concatenate([X.T @ r, [r.sum()]]), "loss": loss, "n": int(X.shape[0])}


# ------------------------------------------------------------------ data-plane flavour
def master_fused(client, data, iterations: int = 20, lr: float = 1.0, rows_per_node: int = 125_000,
                 features: int = 256, organization_ids=None, synthetic: bool = False) -> Dict[str, Any]:
    ids = sorted(organization_ids or [o["id"] for o in client.get_organizations_in_my_collaboration()])
    rv = {"addr": "127.0.0.1", "port": _free_port(), "world": len(ids), "ranks": {str(o): r for r, o in enumerate(ids)}}
    task = client.create_new_task(input_={"method": "fit", "kwargs": dict(iterations=iterations, lr=lr, rendezvous=rv,
                                                                         rows_per_node=rows_per_node, features=features,
                                                                         synthetic=synthetic)},
                                  organization_ids=ids, name="glm-fit")
    client.wait_for_task(task["id"])
    res = sorted(client.get_results(task_id=task["id"]), key=lambda r: r["rank"])
    return {"coefficients": res[0]["coefficients"], "intercept": res[0]["intercept"], "losses": res[0]["losses"],
            "us_per_iteration_max": max(r["us_per_iteration"] for r in res), "world": len(ids),
            "coef_error_vs_truth": res[0].get("coef_error_vs_truth")}


def RPC_fit(data, iterations: int = 20, lr: float = 1.0, rendezvous: Optional[dict] = None,
            rows_per_node: int = 125_000, features: int = 256, synthetic: bool = False) -> Dict[str, Any]:
    import torch
    import torch.distributed as dist

    from ...models.glm import FederatedGLM, synthetic_glm_shard

    org_id = int(os.environ.get("V6_ORGANIZATION_ID", "0"))
    rv = rendezvous or {"addr": "127.0.0.1", "port": _free_port(), "world": 1, "ranks": {str(org_id): 0}}
    rank, world = int(rv["ranks"][str(org_id)]), int(rv["world"])
    use_cuda = torch.cuda.is_available()
    device = torch.device("cuda", int(os.environ.get("V6_GPU", "0"))) if use_cuda else torch.device("cpu")
    if use_cuda:
        torch.cuda.set_device(device)        # all GPUs are visible (peers get mapped): make the pinned one current
    created = False
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl" if use_cuda else "gloo", init_method=f"tcp://{rv['addr']}:{rv['port']}",
                                rank=rank, world_size=world, **({"device_id": device} if use_cuda else {}))
        os.environ["MASTER_PORT"] = str(rv["port"])
        created = True
    w_true = None
    if synthetic or isinstance(data, str) or data is None:      # synthetic://... -> generate this node's shard on its GPU
        X, y, w_true = synthetic_glm_shard(rows_per_node, features, seed=100 + org_id, device=device,
                                           dtype=torch.bfloat16 if use_cuda else torch.float32)
    else:
        Xn, yn = _xy(data)
        X = torch.tensor(Xn, dtype=torch.float32, device=device)
        y = torch.tensor(yn, dtype=torch.float32, device=device)
    glm = FederatedGLM(X, y, rank, world, lr)
    losses = [float(glm.step().item())]
    if use_cuda:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iterations - 1):
        glm.step()
    if use_cuda:
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    losses.append(float(glm.last_loss.item()))
    w = glm.w.detach().cpu().numpy()
    out = {"rank": rank, "coefficients": w[:-1], "intercept": float(w[-1]), "losses": losses,
           "us_per_iteration": 1e6 * dt / max(1, iterations - 1)}
    if w_true is not None:
        out["coef_error_vs_truth"] = float((glm.w - w_true).abs().max().item())
    glm.close()
    if created:
        dist.barrier()
        dist.destroy_process_group()
    return out
