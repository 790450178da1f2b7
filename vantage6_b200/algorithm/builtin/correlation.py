"""Federated covariance / Pearson correlation matrix of numeric columns: every node returns, over its complete cases, the
count, the column sums and the cross-product matrix X'X; they add up across nodes and the master forms the pooled
covariance and correlation -- one round, (p + 1)(p + 2) / 2 numbers per node.  A node with fewer than ``min_rows`` complete
cases refuses (the sums of a handful of rows say too much about them)."""
import numpy as np

from ._common import collect, guard_rows

MIN_ROWS = 10


def master(client, data, columns, organization_ids=None, min_rows: int = MIN_ROWS):
    ids = organization_ids or [o.get("id") for o in client.get_organizations_in_my_collaboration()]
    task = client.create_new_task(input_={"method": "moments", "kwargs": {"columns": list(columns), "min_rows": min_rows}}, organization_ids=ids)
    parts = collect(client, task, ids)
    n = sum(int(p["n"]) for p in parts)
    s = sum(np.asarray(p["sum"], dtype=np.float64) for p in parts)
    xx = sum(np.asarray(p["cross"], dtype=np.float64) for p in parts)
    mean = s / n
    cov = (xx - n * np.outer(mean, mean)) / (n - 1)
    sd = np.sqrt(np.diag(cov))
    with np.errstate(divide="ignore", invalid="ignore"):
        corr = cov / np.outer(sd, sd)
    return {"columns": list(columns), "n": n, "mean": mean, "covariance": cov, "correlation": corr, "n_nodes": len(parts)}


def RPC_moments(data, columns, min_rows: int = MIN_ROWS):
    x = data[list(columns)].dropna().to_numpy(dtype=np.float64)
    guard_rows(x.shape[0], min_rows, "report moments (complete cases)")
    return {"n": int(x.shape[0]), "sum": x.sum(axis=0), "cross": x.T @ x}
