"""Federated Cox proportional-hazards regression (the ``v6-coxph-py`` class of algorithm; WebDISCO formulation, Breslow ties).

Round one pools the distinct event times (optionally binned, as in ``kaplan_meier``) and the per-time sums of the covariates
of the subjects that had the event.  Every Newton iteration then asks each node, for the current coefficients and every
pooled event time t, for three sums over ITS subjects still at risk at t:

    s0(t) = sum exp(x'b)        s1(t) = sum x exp(x'b)        s2(t) = sum x x' exp(x'b)

which add up across nodes; with them the master forms the partial log-likelihood, its gradient and Hessian exactly as a
pooled analysis would, and takes the Newton step.  What leaves a node are sums over risk sets -- no rows.
"""
import numpy as np

from ._common import collect, guard_rows

MIN_ROWS = 10


def _frame(data, time_column, censor_column, columns):
    cols = list(columns) if columns else [c for c in data.columns if c not in (time_column, censor_column)]
    X = data[cols].to_numpy(dtype=np.float64)
    return X, np.asarray(data[time_column], dtype=np.float64), np.asarray(data[censor_column]).astype(bool), cols


def _binned(t, bin_width):
    return np.ceil(t / bin_width) * bin_width if bin_width else t


def master(client, data, time_column: str, censor_column: str, columns=None, organization_ids=None, bin_width: float = 0.0,
           max_iterations: int = 25, tol: float = 1e-9, min_rows: int = MIN_ROWS):
    """``censor_column``: 1 = event observed, 0 = censored.  Returns coefficients, hazard ratios, standard errors, z and the
    partial log-likelihood path."""
    ids = organization_ids or [o.get("id") for o in client.get_organizations_in_my_collaboration()]
    kw = {"time_column": time_column, "censor_column": censor_column, "columns": columns, "bin_width": bin_width, "min_rows": min_rows}
    parts = collect(client, client.create_new_task(input_={"method": "event_sums", "kwargs": kw}, organization_ids=ids), ids)
    names = parts[0]["columns"]
    if any(p["columns"] != names for p in parts):
        raise ValueError("the nodes disagree about the covariates: name them with columns=[...]")
    grid = sorted({float(t) for p in parts for t in p["times"]})
    index = {t: i for i, t in enumerate(grid)}
    p_ = len(names)
    d = np.zeros(len(grid))                      # events per time
    sx = np.zeros((len(grid), p_))               # covariate sums over the events per time
    for part in parts:
        for t, n_ev, s in zip(part["times"], part["events"], part["sum_x"]):
            d[index[float(t)]] += n_ev
            sx[index[float(t)]] += np.asarray(s, dtype=np.float64)
    beta, path = np.zeros(p_), []
    for it in range(1, max_iterations + 1):
        task = client.create_new_task(input_={"method": "risk_sums", "kwargs": {**kw, "grid": grid, "beta": beta.tolist()}},
                                      organization_ids=ids)
        sums = collect(client, task, ids)
        s0 = sum(np.asarray(s["s0"], dtype=np.float64) for s in sums)
        s1 = sum(np.asarray(s["s1"], dtype=np.float64) for s in sums)
        s2 = sum(np.asarray(s["s2"], dtype=np.float64) for s in sums)
        mean = s1 / s0[:, None]
        loglik = float(np.sum(sx @ beta) - np.sum(d * np.log(s0)))
        grad = (sx - d[:, None] * mean).sum(axis=0)
        hess = np.einsum("t,tij->ij", d, s2 / s0[:, None, None] - mean[:, :, None] * mean[:, None, :])
        path.append(loglik)
        step = np.linalg.solve(hess, grad)
        beta = beta + step
        if np.max(np.abs(step)) < tol or (len(path) > 1 and abs(path[-1] - path[-2]) < tol * (abs(path[-1]) + 1)):
            break
    se = np.sqrt(np.diag(np.linalg.inv(hess)))
    return {"columns": names, "coefficients": beta, "hazard_ratios": np.exp(beta), "std_errors": se, "z": beta / se,
            "log_likelihood": path, "iterations": it, "n": int(sum(p["n"] for p in parts)), "n_events": int(d.sum()), "n_nodes": len(parts)}


def RPC_event_sums(data, time_column: str, censor_column: str, columns=None, bin_width: float = 0.0, min_rows: int = MIN_ROWS):
    guard_rows(len(data), min_rows, "take part")
    X, t, ev, cols = _frame(data, time_column, censor_column, columns)
    t = _binned(t, bin_width)
    times = sorted(set(t[ev].tolist()))
    return {"columns": cols, "times": times, "events": [int(((t == u) & ev).sum()) for u in times],
            "sum_x": [X[(t == u) & ev].sum(axis=0) for u in times], "n": int(len(t))}


def RPC_risk_sums(data, time_column: str, censor_column: str, grid, beta, columns=None, bin_width: float = 0.0, min_rows: int = MIN_ROWS):
    guard_rows(len(data), min_rows, "take part")
    X, t, _, _ = _frame(data, time_column, censor_column, columns)
    t = _binned(t, bin_width)
    r = np.exp(X @ np.asarray(beta, dtype=np.float64))
    at_risk = (t[None, :] >= np.asarray(grid, dtype=np.float64)[:, None]).astype(np.float64)          # [times, rows]
    xr = X * r[:, None]
    return {"s0": at_risk @ r, "s1": at_risk @ xr, "s2": np.einsum("tn,ni,nj->tij", at_risk, xr, X)}
