"""Federated weighted mean of a parameter vector (BASELINE.json config 1: "federated
weighted-mean of a 1k-param vector, 2 CPU nodes + 1 server process").

Every node owns ``n_i`` local observations of a d-dimensional vector (rows of its database);
the partial returns ``(sum over rows, n_i)`` and the master forms ``sum_i sum_i / sum_i n_i``
-- i.e. exactly the FedAvg combination rule ``sum_i (n_i / n) * mean_i`` on the control plane.
"""

import numpy as np

from ._common import collect


def _as_matrix(data):
    if isinstance(data, dict):
        data = data.get("x", data.get("data"))
    if hasattr(data, "detach"):            # torch tensor
        data = data.detach().cpu().numpy()
    if hasattr(data, "to_numpy"):           # pandas
        data = data.to_numpy()
    a = np.asarray(data, dtype=np.float64)
    return a.reshape(1, -1) if a.ndim == 1 else a


def master(client, data, organization_ids=None):
    ids = organization_ids or [o.get("id") for o in client.get_organizations_in_my_collaboration()]
    task = client.create_new_task(input_={"method": "partial_sum"}, organization_ids=ids)
    results = collect(client, task, ids)
    total = sum(int(r["count"]) for r in results)
    acc = sum(np.asarray(r["sum"], dtype=np.float64) for r in results)
    return {"mean": acc / total, "count": total, "n_nodes": len(results)}


def RPC_partial_sum(data):
    a = _as_matrix(data)
    return {"sum": a.sum(axis=0), "count": int(a.shape[0])}
