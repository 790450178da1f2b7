"""Federated Kaplan-Meier survival curve (the ``v6-kaplan-meier-py`` class of algorithm).

Round one collects the distinct event times of every node (optionally binned to ``bin_width`` so that single patients
cannot be told apart by their exact time); round two asks every node, for the pooled time grid, how many subjects were at
risk and how many had the event at each time.  The master multiplies ``1 - d_t / n_t`` along the grid -- the estimator a
pooled analysis would give, without any row leaving a node.
"""
import math

import numpy as np

from ._common import collect, guard_rows

MIN_ROWS = 10


def _times(data, time_column, bin_width):
    t = np.asarray(data[time_column], dtype=np.float64)
    return np.ceil(t / bin_width) * bin_width if bin_width else t


def master(client, data, time_column: str, censor_column: str, organization_ids=None, bin_width: float = 0.0, min_rows: int = MIN_ROWS):
    """``censor_column``: 1 = event observed, 0 = censored."""
    ids = organization_ids or [o.get("id") for o in client.get_organizations_in_my_collaboration()]
    kw = {"time_column": time_column, "bin_width": bin_width, "min_rows": min_rows}
    t = client.create_new_task(input_={"method": "event_times", "kwargs": {**kw, "censor_column": censor_column}}, organization_ids=ids)
    grid = sorted({float(x) for p in collect(client, t, ids) for x in p["times"]})
    t2 = client.create_new_task(input_={"method": "risk_table", "kwargs": {**kw, "censor_column": censor_column, "grid": grid}},
                                organization_ids=ids)
    parts = collect(client, t2, ids)
    at_risk = np.sum([np.asarray(p["at_risk"], dtype=np.float64) for p in parts], axis=0)
    events = np.sum([np.asarray(p["events"], dtype=np.float64) for p in parts], axis=0)
    surv, var_acc, s, curve = 1.0, 0.0, [], []
    for ti, n, d in zip(grid, at_risk, events):
        if n > 0:
            surv *= 1.0 - d / n
            if n > d:
                var_acc += d / (n * (n - d))            # Greenwood
        se = surv * math.sqrt(var_acc)
        curve.append({"time": ti, "at_risk": int(n), "events": int(d), "survival": surv, "std_err": se})
        s.append(surv)
    below = [c["time"] for c in curve if c["survival"] <= 0.5]
    return {"curve": curve, "median_survival": below[0] if below else None, "n": int(sum(p["n"] for p in parts)), "n_nodes": len(parts)}


def RPC_event_times(data, time_column: str, censor_column: str, bin_width: float = 0.0, min_rows: int = MIN_ROWS):
    guard_rows(len(data), min_rows, "report event times")
    t = _times(data, time_column, bin_width)
    ev = np.asarray(data[censor_column]).astype(bool)
    return {"times": sorted(set(t[ev].tolist()))}


def RPC_risk_table(data, time_column: str, censor_column: str, grid, bin_width: float = 0.0, min_rows: int = MIN_ROWS):
    guard_rows(len(data), min_rows, "report a risk table")
    t = _times(data, time_column, bin_width)
    ev = np.asarray(data[censor_column]).astype(bool)
    g = np.asarray(grid, dtype=np.float64)
    at_risk = (t[None, :] >= g[:, None]).sum(axis=1)
    events = ((t[None, :] == g[:, None]) & ev[None, :]).sum(axis=1)
    return {"at_risk": at_risk.tolist(), "events": events.tolist(), "n": int(len(t))}
