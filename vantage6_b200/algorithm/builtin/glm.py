"""Federated GLM as a vantage6 algorithm (BASELINE config 5: logistic regression).

Three flavours:
* ``master_irls`` / ``RPC_irls_partial`` -- Fisher scoring on the control plane for the gaussian, binomial (logit) and
  poisson (log) families, the ``v6-glm-py`` formulation: each iteration every node returns its (p+1)x(p+1) ``X'WX`` and
  ``X'Wz``, the master solves the pooled normal equations -- the coefficients, standard errors and deviance a pooled fit
  would give, in a handful of round trips;
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

from ._common import collect, guard_rows
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
        res = collect(client, task, ids)
        g = sum(np.asarray(r["grad"], dtype=np.float64) for r in res)
        n = sum(r["n"] for r in res)
        if w is None:
            w = np.zeros_like(g)
        w = w - lr * g / n
        losses.append(sum(r["loss"] for r in res) / n)
    return {"coefficients": w[:-1], "intercept": float(w[-1]), "losses": losses, "n": int(n)}


def RPC_gradient(data, w=None) -> Dict[str, Any]:
    X, y = _xy(data)
    guard_rows(X.shape[0], None, "report a gradient")          # the gradient of a handful of rows gives them away
    wv = np.zeros(X.shape[1] + 1) if w is None else np.asarray(w, dtype=np.float64)
    z = X @ wv[:-1] + wv[-1]
    p = 1.0 / (1.0 + np.exp(-z))
    r = p - y
    loss = float(np.sum(np.maximum(z, 0) - z * y + np.log1p(np.exp(-np.abs(z)))))
    return {"grad": np.concatenate([X.T @ r, [r.sum()]]), "loss": loss, "n": int(X.shape[0])}


# ------------------------------------------------------------------ Fisher scoring (IRLS) for three families
FAMILIES = ("gaussian", "binomial", "poisson")


def _design(data, columns=None, outcome=None):
    """(X with a trailing intercept column, y).  Data frames may name the ``outcome`` column and the feature ``columns``;
    otherwise the last column is the outcome."""
    if outcome is not None and hasattr(data, "columns"):
        feats = list(columns) if columns else [c for c in data.columns if c != outcome]
        X, y = data[feats].to_numpy(dtype=np.float64), data[outcome].to_numpy(dtype=np.float64)
    else:
        X, y = _xy(data)
    return np.hstack([X, np.ones((X.shape[0], 1))]), y


def _mean_and_weights(family: str, eta, y):
    """Mean, IRLS weights, working response and the unit deviance for the canonical link of ``family``."""
    if family == "gaussian":
        mu = eta
        return mu, np.ones_like(eta), y, (y - mu) ** 2
    if family == "binomial":
        mu = np.clip(1.0 / (1.0 + np.exp(-eta)), 1e-10, 1 - 1e-10)
        var = mu * (1 - mu)
        with np.errstate(divide="ignore", invalid="ignore"):
            dev = 2 * (np.where(y > 0, y * np.log(y / mu), 0.0) + np.where(y < 1, (1 - y) * np.log((1 - y) / (1 - mu)), 0.0))
        return mu, var, eta + (y - mu) / var, dev
    if family == "poisson":
        mu = np.exp(np.clip(eta, -30, 30))
        with np.errstate(divide="ignore", invalid="ignore"):
            dev = 2 * (np.where(y > 0, y * np.log(y / mu), 0.0) - (y - mu))
        return mu, mu, eta + (y - mu) / mu, dev
    raise ValueError(f"unknown family {family!r} (known: {FAMILIES})")


def master_irls(client, data, family: str = "binomial", columns=None, outcome=None, max_iterations: int = 25, tol: float = 1e-8,
                organization_ids=None) -> Dict[str, Any]:
    if family not in FAMILIES:
        raise ValueError(f"unknown family {family!r} (known: {FAMILIES})")
    ids = organization_ids or [o["id"] for o in client.get_organizations_in_my_collaboration()]
    beta, deviance, history = None, None, []
    for it in range(1, max_iterations + 1):
        kw = {"family": family, "columns": columns, "outcome": outcome, "beta": None if beta is None else beta.tolist()}
        task = client.create_new_task(input_={"method": "irls_partial", "kwargs": kw}, organization_ids=ids)
        parts = collect(client, task, ids)
        xtwx = sum(np.asarray(p["xtwx"], dtype=np.float64) for p in parts)
        xtwz = sum(np.asarray(p["xtwz"], dtype=np.float64) for p in parts)
        n = sum(int(p["n"]) for p in parts)
        new_dev = float(sum(p["deviance"] for p in parts))            # deviance AT the beta that was sent out
        if beta is not None:                                            # (the first pass starts from the data, not from a model)
            history.append(new_dev)
        new_beta = np.linalg.solve(xtwx, xtwz)
        converged = beta is not None and (abs(new_dev - deviance) <= tol * (abs(new_dev) + 0.1) or family == "gaussian")
        beta, deviance = new_beta, new_dev
        if converged:
            break
    p_ = beta.shape[0]
    dispersion = deviance / max(1, n - p_) if family == "gaussian" else 1.0
    cov = np.linalg.inv(xtwx) * dispersion
    return {"family": family, "coefficients": beta[:-1], "intercept": float(beta[-1]), "std_errors": np.sqrt(np.diag(cov)),
            "deviance": deviance, "deviance_history": history, "dispersion": dispersion, "iterations": it, "n": n, "n_nodes": len(parts)}


def RPC_irls_partial(data, family: str = "binomial", columns=None, outcome=None, beta=None) -> Dict[str, Any]:
    X, y = _design(data, columns, outcome)
    guard_rows(X.shape[0], None, "report normal equations")
    if beta is None:                                   # start from the family's usual initial mean
        if family == "gaussian":
            eta = y
        elif family == "binomial":
            mu0 = (y + 0.5) / 2
            eta = np.log(mu0 / (1 - mu0))
        else:
            eta = np.log(np.maximum(y, 0) + 0.1)
    else:
        eta = X @ np.asarray(beta, dtype=np.float64)
    _, w, z, dev = _mean_and_weights(family, eta, y)
    xw = X * w[:, None]
    return {"xtwx": xw.T @ X, "xtwz": xw.T @ z, "deviance": float(dev.sum()), "n": int(X.shape[0])}


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
